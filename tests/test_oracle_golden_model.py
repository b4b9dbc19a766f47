"""The oracle against its own frozen outputs on the scaled-down synthetic model (tests/golden/tiny_model_fixture.json,
made by tests/golden/make_golden.py).  The reference cannot run here, so these vectors do not pin the oracle to the
reference -- they pin it in time: an edit under oracle/ or a torch upgrade that moves a code, a logit or a PCM sample
fails here before it silently moves the target of the GPU parity tests."""
import json
import os

import numpy as np
import torch

from oracle import generate as OG, model as OM, sampling as osmp, vocoder as OV
from qwen3_tts_rs_b200 import spec as S, weights as W
from conftest import talker_weights, vocoder_weights

FIX = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_model_fixture.json")))


def test_prompt_and_weights_are_the_fixture_ones():
    spec = S.SPECS[FIX["spec"]]
    assert W.synthetic_prompt(FIX["prompt_index"], spec) == FIX["text_ids"]


def test_generation_matches_the_frozen_codes_and_logits():
    spec = S.SPECS[FIX["spec"]]
    tw = talker_weights(spec)
    for mode, prec in (("bf16", OM.BF16P), ("f32", OM.F32P)):
        g = FIX[mode]
        tk, cp = OM.Talker(spec, tw, prec), OM.CodePredictor(spec, tw, prec)
        emb = tk.custom_voice_embeds(FIX["text_ids"], S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
        frames = OG.prefill_and_generate(tk, cp, emb, FIX["text_ids"], osmp.GenerationConfig(max_new_tokens=FIX["frames"]), FIX["seed"])
        assert frames == g["codes"], mode
        hidden, logits = tk.run_prefill_layers(emb, tk.new_kv_caches())
        top = torch.topk(logits[0, 0].float(), 5)
        assert top.indices.tolist() == g["prefill_top5_ids"]
        assert np.allclose([float(v) for v in top.values], g["prefill_top5_logits"], rtol=1e-5, atol=1e-6)
        assert abs(float(hidden[0, -1].float().norm()) - g["last_hidden_l2"]) <= 1e-4 * g["last_hidden_l2"]
    # the two precisions agree on this utterance's first frame (a sanity link between the F32 CPU path of configs[0]
    # and the bf16 path the CUDA kernels follow)
    assert FIX["bf16"]["codes"][0][0] == FIX["f32"]["codes"][0][0]


def test_vocoder_matches_the_frozen_waveform():
    spec = S.SPECS[FIX["spec"]]
    v = FIX["vocoder"]
    pcm = OV.Vocoder(spec.vocoder, vocoder_weights(spec.vocoder, spec.name)).decode(OG.codes_to_tensor(v["codes"]))[0, 0].numpy()
    assert pcm.size == v["n_samples"] == len(v["codes"]) * 1920
    assert abs(float(np.sqrt(np.mean(pcm.astype(np.float64) ** 2))) - v["rms"]) <= 1e-5 * v["rms"]
    for got, want in ((pcm[:32], v["first32"]), (pcm[-32:], v["last32"]), (pcm[::97], v["every_97th"])):
        assert np.abs(got - np.asarray(want, dtype=np.float32)).max() <= 2e-6
