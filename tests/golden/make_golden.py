"""Generates tests/golden/*.json from the oracle (the reference cannot run here: no cargo, no weights).
The fixtures freeze the oracle's outputs on the reference's own weight-free sampler fixture
(benches/sampling.rs:12-64) so that later oracle edits cannot silently change them.
Run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import sampling as s  # noqa: E402

F = np.float32
i = np.arange(3072, dtype=F)
logits = (np.sin(i * F(0.1)) * F(5.0)).astype(F)[None]
cases = []
for top_k, top_p in [(50, 1.0), (0, 0.5), (0, 0.9), (0, 0.95), (50, 0.9)]:
    cfg = s.GenerationConfig(temperature=0.9, top_k=top_k, top_p=top_p, repetition_penalty=1.0)
    ctx = s.SamplingContext(42)
    cases.append(dict(top_k=top_k, top_p=top_p, tokens=[int(s.sample(logits, cfg, ctx)[0]) for _ in range(32)]))
ctx = s.SamplingContext(42)
json.dump(dict(cases=cases, pcg_seed42_u32=[ctx.next_u32() for _ in range(8)]),
          open(os.path.join(HERE, "sampler_fixture.json"), "w"), indent=1)
print("wrote sampler_fixture.json")


# ---- model-path fixture: the oracle's outputs on the scaled-down synthetic model, frozen ---------------------------------
# (the reference itself cannot produce these here; they pin the ORACLE so that an edit to oracle/ or a torch upgrade cannot
# move the target the CUDA path is compared with, and they give the GPU tests a committed vector to check against)
def model_fixture():
    import torch
    from oracle import generate as OG, model as OM, vocoder as OV
    from qwen3_tts_rs_b200 import spec as S, weights as W
    spec = S.SPEC_TINY
    tw, vw = W.make_talker_weights(spec), W.make_vocoder_weights(spec.vocoder)
    ids = W.synthetic_prompt(0, spec)
    out = dict(spec=spec.name, prompt_index=0, text_ids=ids, seed=42, frames=6)
    for mode, prec in (("bf16", OM.BF16P), ("f32", OM.F32P)):
        tk, cp = OM.Talker(spec, tw, prec), OM.CodePredictor(spec, tw, prec)
        emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
        tr = OG.Trace()
        frames = OG.prefill_and_generate(tk, cp, emb, ids, s.GenerationConfig(max_new_tokens=6), 42, trace=tr)
        hidden, logits = tk.run_prefill_layers(emb, tk.new_kv_caches())
        top = torch.topk(logits[0, 0].float(), 5)
        out[mode] = dict(codes=frames, prefill_top5_ids=top.indices.tolist(), prefill_top5_logits=[float(v) for v in top.values],
                         prefill_logits_sum=float(logits.float().sum()), last_hidden_l2=float(hidden[0, -1].float().norm()))
    codes = [[(f * 37 + q * 101 + 5) % 2048 for q in range(16)] for f in range(5)]
    pcm = OV.Vocoder(spec.vocoder, vw).decode(OG.codes_to_tensor(codes))[0, 0].numpy()
    out["vocoder"] = dict(codes=codes, n_samples=int(pcm.size), rms=float(np.sqrt(np.mean(pcm.astype(np.float64) ** 2))),
                          first32=[float(v) for v in pcm[:32]], last32=[float(v) for v in pcm[-32:]],
                          every_97th=[float(v) for v in pcm[::97]])
    return out


json.dump(model_fixture(), open(os.path.join(HERE, "tiny_model_fixture.json"), "w"), indent=1)
print("wrote tiny_model_fixture.json")
