"""Run a few decode frames of the bench workload (for ncu).  Prints the library's kernel-launch counter
before the frame loop so that `ncu -s <count>` can skip model loading and prefill."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="1.7b")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--vocoder", action="store_true")
a = ap.parse_args()
spec = S.SPECS[a.model]
tw = W.make_talker_weights(spec)
vw = W.make_vocoder_weights(spec.vocoder) if a.vocoder else None
tts = api.Qwen3TTS.from_weights(spec, tw, vw)
prompts = [W.synthetic_prompt(i, spec) for i in range(a.batch)]
pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
opts = api.SynthesisOptions(max_length=max(a.frames, 8))
sess = api.Session(tts.model, a.batch, opts, [42 + i for i in range(a.batch)], max_seq=512)
sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp])
sess.set_trailing_ids([list(t[1:]) for t in prompts])
sess.synchronize()
print("LAUNCHES_BEFORE_LOOP", L.load().q3_kernel_launch_count(), flush=True)
codes, n = sess.generate(a.frames)
print("LAUNCHES_AFTER_LOOP", L.load().q3_kernel_launch_count(), "frames", n.tolist(), flush=True)
if a.vocoder:
    sess.vocode(a.frames, to_host=False)
    sess.synchronize()
    print("LAUNCHES_AFTER_VOCODER", L.load().q3_kernel_launch_count(), flush=True)
