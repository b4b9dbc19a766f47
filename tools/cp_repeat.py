"""Race hunt: the code-predictor frame (q3_code_predictor_frame, one launch of the persistent kernel) repeated with
identical inputs; any difference in the logits between repetitions is a race.  Reports, per repetition that differs, the
first pass whose logits differ and how many rows / logits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qwen3_tts_rs_b200 import api, spec as S, weights as W
name = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 200
spec = S.SPECS[name]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=64), list(range(B)), max_seq=128)
g = torch.Generator().manual_seed(5)
hid = (torch.randn(B, spec.hidden, generator=g) * 0.7).to(torch.bfloat16)
toks = [100 + 37 * i for i in range(B)]
ref_codes, ref = sess.code_predictor_frame(hid, toks, want_logits=True)
bad = {}
for it in range(N):
    codes, lg = sess.code_predictor_frame(hid, toks, want_logits=True)
    if not np.array_equal(lg, ref):
        d = np.argwhere(lg != ref)          # [b, g, v]
        g0 = int(d[:, 1].min())
        rows = sorted(set(int(x[0]) for x in d if x[1] == g0))
        bad.setdefault(g0, []).append((it, rows, int((lg[:, g0] != ref[:, g0]).sum()), float(np.abs(lg[:, g0] - ref[:, g0]).max())))
print(f"{name} B={B} Q3_MEGA={os.environ.get('Q3_MEGA','default')} slots={os.environ.get('Q3_M4_SLOTS','max')} flags={os.environ.get('Q3_PF_SLEEP','0')}: "
      f"{sum(len(v) for v in bad.values())}/{N} repetitions differ; first differing pass -> count: { {k: len(v) for k, v in sorted(bad.items())} }")
for k in sorted(bad)[:3]:
    print("   pass", k, "examples (iteration, rows, n logits, max |d|):", bad[k][:4])
