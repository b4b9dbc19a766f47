"""Timestamp breakdown of the persistent frame kernel (block 0): stage / tiles / barrier per phase."""
import os, sys, ctypes as C
os.environ.setdefault("Q3TTS_LIB", "dev")      # profiling hooks live in the development library
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="1.7b"); ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=3); a = ap.parse_args()
spec = S.SPECS[a.model]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
prompts = [W.synthetic_prompt(i, spec) for i in range(a.batch)]
pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
sess = api.Session(tts.model, a.batch, api.SynthesisOptions(max_length=64), [42 + i for i in range(a.batch)], max_seq=512)
sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp]); sess.set_trailing_ids([list(t[1:]) for t in prompts])
sess.generate(4)   # warm
lib0 = L.load()
lib0.q3_debug_barrier_bench.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
ms = C.c_float(0)
L.check(lib0.q3_debug_barrier_bench(sess.handle, 2000, C.byref(ms)))
print(f"grid barrier: {ms.value * 1e3 / 2000:.2f} us each (2000 back-to-back, BAR_MODE={os.environ.get('Q3_BAR_MODE','0')})")
import time
sess.synchronize(); t0 = time.perf_counter(); sess.generate(36); dt = time.perf_counter() - t0
print(f"32 frames: {dt*1e3/32:.3f} ms/frame (PREFETCH={os.environ.get('Q3_PREFETCH','2')})")
lib = L.load()
cap = 20000
buf = np.zeros(cap, dtype=np.uint64); n = C.c_int32(0)
lib.q3_debug_profile.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
L.check(lib.q3_debug_profile(sess.handle, a.frames, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
st = buf[: n.value]
t = (st >> np.uint64(8)).astype(np.int64); tag = (st & np.uint64(0xff)).astype(int)
print("stamps", n.value, "total us", (t[-1] - t[0]) / 1e3)
# tags: 1 gemv entry, 2 loads issued.., 6 first loads + scales done, 7 tile loop done, 8 combine barrier passed, 3 phase end, 4/5 attention
seg = {}
cnt = {}
phases = []   # per gemv phase: dict of segment ns
cur = None
for i in range(1, len(t)):
    d = int(t[i] - t[i - 1]); k = (int(tag[i - 1]), int(tag[i]))
    seg[k] = seg.get(k, 0) + d; cnt[k] = cnt.get(k, 0) + 1
    if tag[i - 1] == 1: cur = {}
    if cur is not None:
        cur[k] = d
        if tag[i] == 3: phases.append(cur); cur = None
for k in sorted(seg): print(f"  {k}: n={cnt[k]:5d} total {seg[k]/1e3:9.1f} us  mean {seg[k]/cnt[k]:7.0f} ns")
per_frame = len(phases) // a.frames
print("gemv phases per frame", per_frame)
names = ["proj"] + ["qkv", "o", "gateup", "down"] * 5 + ["head"]
def show(label, ps, nm):
    print(label)
    for n_, p_ in zip(nm, ps):
        print(f"   {n_:7s} load+scale {p_.get((2,6),0):6d}  tiles {p_.get((6,7),0):6d}  sync {p_.get((7,8),0):5d}  combine+epi {p_.get((8,3),0):6d}")
f0 = phases[per_frame:2 * per_frame] if a.frames > 1 else phases[:per_frame]
show("CP pass 0", f0[:22], names)
show("CP pass 1", f0[22:44], names)
show("CP pass 7", f0[22 * 7:22 * 8], names)
show("talker layer 0-1", f0[-113:-105], ["qkv", "o", "gateup", "down"] * 2)
show("talker last + head", f0[-5:], ["qkv", "o", "gateup", "down", "head"])
