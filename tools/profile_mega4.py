"""Timestamp breakdown of the TMA-ring frame kernel (block 0, mega4.cuh); also works for mega2.cuh (no tags 8 / 10 there).
stamps: 8 phase top (descriptor ready), 1 gemv entry, 2 wait passed, 3 activations loaded (+ scales), 4 last tile's MMAs done +
combine barrier, 5 gemv end, 6 attention wait passed, 7 attention end; tag 10 carries a DURATION: ns thread 0 spent waiting for
ring slots in the phase."""
import os, sys, ctypes as C
os.environ.setdefault("Q3TTS_LIB", "dev")      # profiling hooks live in the development library
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W
import argparse, time
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="1.7b"); ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=3); a = ap.parse_args()
spec = S.SPECS[a.model]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
prompts = [W.synthetic_prompt(i, spec) for i in range(a.batch)]
pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
sess = api.Session(tts.model, a.batch, api.SynthesisOptions(max_length=400), [42 + i for i in range(a.batch)], max_seq=512)
sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp]); sess.set_trailing_ids([list(t[1:]) for t in prompts])
sess.generate(4)   # warm
lib = L.load()
import torch
stream = torch.cuda.ExternalStream(lib.q3_session_stream(sess.handle))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sess.synchronize(); e0.record(stream); sess.generate_async(68); e1.record(stream); sess.synchronize()
print(f"64 frames: {e0.elapsed_time(e1)/64:.3f} ms/frame (Q3_MEGA={os.environ.get('Q3_MEGA','default')} slots={os.environ.get('Q3_M4_SLOTS','max')})")
cap = 40000
buf = np.zeros(cap, dtype=np.uint64); n = C.c_int32(0)
lib.q3_debug_profile.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
L.check(lib.q3_debug_profile(sess.handle, a.frames, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
st = buf[: n.value]
val = (st >> np.uint64(8)).astype(np.int64); tag = (st & np.uint64(0xff)).astype(int)
# split duration records (tag 10) from time stamps
phases, cur, slotw = [], None, 0
seg, cnt = {}, {}
prev_t, prev_tag = None, None
for v, tg in zip(val, tag):
    if tg == 10:
        slotw = int(v); continue
    if prev_t is not None:
        k = (prev_tag, int(tg)); d = int(v - prev_t)
        seg[k] = seg.get(k, 0) + d; cnt[k] = cnt.get(k, 0) + 1
        if prev_tag == 1: cur = {}
        if cur is not None:
            cur[k] = d
            if tg == 5:
                cur["slotw"] = slotw; phases.append(cur); cur = None
    prev_t, prev_tag = int(v), int(tg)
tt = val[tag != 10]
print("stamps", n.value, "per frame us", (tt[-1] - tt[0]) / 1e3 / a.frames)
for k in sorted(seg): print(f"  {k}: n={cnt[k]:5d} total {seg[k]/1e3/a.frames:9.1f} us/frame  mean {seg[k]/cnt[k]:7.0f} ns")
per_frame = len(phases) // a.frames
print("gemv phases per frame", per_frame, " slot-wait total us/frame", sum(p_["slotw"] for p_ in phases) / 1e3 / a.frames)
names = ["proj"] + ["qkv", "o", "gateup", "down"] * 5 + ["head"]
def show(label, ps, nm):
    print(label)
    for n_, p_ in zip(nm, ps):
        print(f"   {n_:7s} wait {p_.get((1,2),0):6d}  x+scale {p_.get((2,3),0):6d}  mma {p_.get((3,4),0):6d} (slot wait {p_.get('slotw',0):6d})  combine+store {p_.get((4,5),0):6d}")
f0 = phases[per_frame:2 * per_frame] if a.frames > 1 else phases[:per_frame]
show("CP pass 0", f0[:22], names)
show("CP pass 7", f0[22 * 7:22 * 8], names)
show("talker layer 0-1", f0[-113:-105], ["qkv", "o", "gateup", "down"] * 2)
show("talker last + head", f0[-5:], ["qkv", "o", "gateup", "down", "head"])
