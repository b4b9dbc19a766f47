"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): tiny spec, batch 2, 3 frames through the persistent
kernel, the sampler and the vocoder.  racecheck sees shared-memory hazards only; the global tagged-slot protocol is covered by
tests/test_gpu_parity.py::test_code_predictor_frame_repeats_bit_for_bit_at_1p7b."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_tts_rs_b200 import api, spec as S, weights as W
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
spec = S.SPECS[name]
vw = W.make_vocoder_weights(spec.vocoder)
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), vw)
prompts = [W.synthetic_prompt(i, spec) for i in range(2)]
audio = tts.synthesize_with_voice(prompts, options=api.SynthesisOptions(max_length=3), seeds=[1, 2])
print("frames", [len(a) // 1920 for a in audio], "generation", api.Session(tts.model, 2, api.SynthesisOptions(max_length=3), [1, 2]).decode_generation())
