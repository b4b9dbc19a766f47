"""Long-utterance soak: the reference's default frame budget (max_new_tokens 2048, lib.rs SynthesisOptions) at 1.7B dimensions --
2048 frames per row through the persistent kernel (context 10 + 2048 positions, 128 launches of 16 frames), twice (bit-equal
codes), then the vocoder over 2048 frames (3.9 M samples per row).  Reports frames/s and any error."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qwen3_tts_rs_b200 import api, spec as S, weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
F = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
spec = S.SPECS["1.7b"]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), W.make_vocoder_weights(spec.vocoder))
prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
opts = api.SynthesisOptions(max_length=F, eos_token_id=None)        # no early stop: every row runs the whole budget
seeds = [42 + i for i in range(B)]
t0 = time.perf_counter(); a = tts.generate_codes(prompts, options=opts, seeds=seeds); t1 = time.perf_counter()
b = tts.generate_codes(prompts, options=opts, seeds=seeds); t2 = time.perf_counter()
print(f"codes: {[len(r) for r in a]} frames per row, {B * F / (t2 - t1):.0f} frames/s, repeat identical: {a == b}", flush=True)
if os.environ.get("SOAK_CODES_ONLY"): sys.exit(0)
audio = tts.synthesize_with_voice(prompts[:2], options=opts, seeds=seeds[:2])
print("audio samples per row", [len(x) for x in audio], "finite", all(np.isfinite(x.samples).all() for x in audio),
      "peak", [float(np.abs(x.samples).max()) for x in audio])
# the streamed (stateful) form of the same utterance: cache capacity, carried window
st = tts.synthesize_streaming(prompts[0], options=api.SynthesisOptions(max_length=F, eos_token_id=None, seed=42, stream_left_context=-1, stream_first_chunk=2))
n = 0; t0 = time.perf_counter()
for c in st: n += len(c.samples)
print(f"stateful stream: {n // 1920} frames in {time.perf_counter() - t0:.2f} s; equals non-streamed: {n == len(audio[0])}")
