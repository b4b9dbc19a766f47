"""Localise a nondeterminism: tapped runs (per-frame CP logits of every pass, talker logits, talker input) compared with a
saved reference run.  usage: race_localize.py save ref.npz | race_localize.py cmp ref.npz"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from qwen3_tts_rs_b200 import api, spec as S, weights as W
from test_gpu_parity import run_tapped
mode, path = sys.argv[1], sys.argv[2]
B, F = 8, int(os.environ.get("FRAMES", "40"))
spec = S.SPECS["1.7b"]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
seeds = [42 + i for i in range(B)]
opts = api.SynthesisOptions(max_length=F)
def run():
    codes, taps = run_tapped(tts, prompts, seeds, opts, F)
    return dict(codes=np.asarray(codes), cp=taps["cp_logits"], lg=taps["logits"], si=taps["step_input"].view(dtype=__import__("torch").int16).numpy())
if mode == "save":
    r = run(); np.savez(path, **r); print("saved", int(r["codes"].astype("int64").sum()))
else:
    ref = np.load(path)
    for rep in range(int(os.environ.get("REPS", "4"))):
        r = run()
        msg = "identical"
        for f in range(F):
            d = np.argwhere(r["cp"][f] != ref["cp"][f])
            if len(d):
                g, b = int(d[0][0]), int(d[0][1])
                nbad = int((r["cp"][f, g, b] != ref["cp"][f, g, b]).sum())
                maxd = float(np.abs(r["cp"][f, g, b] - ref["cp"][f, g, b]).max())
                rows = sorted(set(int(x[1]) for x in d if x[0] == g))
                msg = f"first diff: frame {f} CP pass {g} rows {rows}: row {b} has {nbad}/{r['cp'].shape[-1]} logits differing, max |d| {maxd:.4f}; step_input of frame {f-1} equal: {bool((r['si'][f-1] == ref['si'][f-1]).all()) if f else None}; talker logits frame {f-1} equal: {bool((r['lg'][f-1] == ref['lg'][f-1]).all()) if f else None}"
                break
            if (r["lg"][f] != ref["lg"][f]).any():
                b = int(np.argwhere(r["lg"][f] != ref["lg"][f])[0][0])
                msg = f"first diff: frame {f} TALKER logits row {b} ({int((r['lg'][f,b] != ref['lg'][f,b]).sum())} differ, max |d| {float(np.abs(r['lg'][f,b]-ref['lg'][f,b]).max()):.4f}); cp logits of the frame equal; step_input equal: {bool((r['si'][f] == ref['si'][f]).all())}"
                break
        print(f"rep {rep}: {msg}", flush=True)
