"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): tiny spec, batch 2, 3 frames through the persistent
kernel, the sampler and the vocoder (tcgen05 convolutions), then the speaker encoder and a voice-clone (ICL) prefill + decode.  racecheck sees shared-memory hazards only; the global tagged-slot protocol is covered by
tests/test_gpu_parity.py::test_code_predictor_frame_repeats_bit_for_bit_at_1p7b."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_tts_rs_b200 import api, spec as S, weights as W
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
spec = S.SPECS[name]
import numpy as np, torch
vw = dict(W.make_vocoder_weights(spec.vocoder))
scfg = S.SpeakerSpec(mel_dim=32, enc_dim=spec.hidden, enc_channels=(64, 64, 64, 64, 192), enc_attention_channels=32,
                     enc_res2net_scale=4, enc_se_channels=32)
vw.update(W.make_speaker_weights(scfg))
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), vw)
prompts = [W.synthetic_prompt(i, spec) for i in range(2)]
audio = tts.synthesize_with_voice(prompts, options=api.SynthesisOptions(max_length=3), seeds=[1, 2])
print("frames", [len(a) // 1920 for a in audio], "generation", api.Session(tts.model, 2, api.SynthesisOptions(max_length=3), [1, 2]).decode_generation())
emb = tts.speaker_encode(np.random.default_rng(0).standard_normal((1, 32, 40)).astype(np.float32))[0]
ref = np.random.default_rng(1).integers(0, 2048, size=(5, 16)).astype(np.uint32)
clone = tts.synthesize_voice_clone([prompts[0][:4]], [api.VoiceClonePrompt(emb, ref, [3, 4, 5])],
                                   options=api.SynthesisOptions(max_length=3, eos_token_id=None), seeds=[1])
print("voice clone samples", [len(a) for a in clone])
