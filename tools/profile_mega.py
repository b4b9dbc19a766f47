"""Timestamp breakdown of the persistent frame kernel (block 0): stage / tiles / barrier per phase."""
import os, sys, ctypes as C
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
# tags: 1 gemv entry, 2 after stage, 3 after tiles, 4 attn start, 5 attn end
stage = tiles = attn = other = 0
cnt = {1: 0, 4: 0}
for i in range(1, len(t)):
    d = t[i] - t[i - 1]
    if tag[i - 1] == 1 and tag[i] == 2: stage += d; cnt[1] += 1
    elif tag[i - 1] == 2 and tag[i] == 3: tiles += d
    elif tag[i - 1] == 4 and tag[i] == 5: attn += d; cnt[4] += 1
    else: other += d
print(f"gemv phases {cnt[1]}: stage {stage/1e3:.1f} us ({stage/max(1,cnt[1]):.0f} ns each), tiles {tiles/1e3:.1f} us ({tiles/max(1,cnt[1]):.0f} ns each)")
print(f"attn phases {cnt[4]}: {attn/1e3:.1f} us ({attn/max(1,cnt[4]):.0f} ns each); barriers+other {other/1e3:.1f} us ({other/max(1,cnt[1]+cnt[4]):.0f} ns per phase)")
# first 40 deltas for a look at the CP pass 0 and talker layer 0
print([(int(tag[i]), int(t[i] - t[i-1])) for i in range(1, min(60, len(t)))])
# --- per-phase detail for one frame: list (stage_ns, tiles_ns) of consecutive gemv phases near the end (talker) and start (CP)
ph = []
i = 0
while i < len(t) - 2:
    if tag[i] == 1 and tag[i + 1] == 2 and tag[i + 2] == 3:
        ph.append((int(t[i + 1] - t[i]), int(t[i + 2] - t[i + 1])))
        i += 3
    else:
        i += 1
per_frame = len(ph) // a.frames
print("gemv phases per frame", per_frame)
f0 = ph[:per_frame]
print("CP pass 0 (proj, [qkv,o,gateup,down]x5, head) tiles ns:", [x[1] for x in f0[:22]])
print("CP pass 1 tiles ns:", [x[1] for x in f0[22:44]])
print("talker layers 0-1 (qkv,o,gateup,down) tiles ns:", [x[1] for x in f0[-113:-105]], " last layer+head:", [x[1] for x in f0[-5:]])
