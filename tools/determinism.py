"""Run the same generation several times (fresh session each) and report whether the codes repeat bit for bit."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qwen3_tts_rs_b200 import api, spec as S, weights as W
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="1.7b"); ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=48); ap.add_argument("--reps", type=int, default=4); a = ap.parse_args()
spec = S.SPECS[a.model]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
prompts = [W.synthetic_prompt(i, spec) for i in range(a.batch)]
seeds = [42 + i for i in range(a.batch)]
opts = api.SynthesisOptions(max_length=a.frames)
runs = [np.asarray(tts.generate_codes(prompts, options=opts, seeds=seeds)) for _ in range(a.reps)]
first = []
for r in runs[1:]:
    d = np.argwhere(r != runs[0])
    first.append(None if len(d) == 0 else tuple(int(x) for x in d[0]))
print(f"Q3_MEGA={os.environ.get('Q3_MEGA','default')} slots={os.environ.get('Q3_M4_SLOTS','max')} batch={a.batch}: checksum {[int(r.astype('int64').sum()) for r in runs]} first diff (row, frame, code) vs run 0: {first}", flush=True)
