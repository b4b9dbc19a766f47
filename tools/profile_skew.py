"""Arrival-time skew of the CTAs at every phase of one frame (dataflow kernel, Q3_PROF_MODE=2)."""
import os, sys, ctypes as C
os.environ["Q3_PROF_MODE"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W
spec = S.SPECS["1.7b"]; B = 8
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=400), [42 + i for i in range(B)], max_seq=512)
sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp]); sess.set_trailing_ids([list(t[1:]) for t in prompts])
sess.generate(4)
lib = L.load()
G = 148; NPH = 549
cap = (NPH * 4 + 8) * G + 4096
buf = np.zeros(cap, dtype=np.uint64); n = C.c_int32(0)
lib.q3_debug_profile.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
L.check(lib.q3_debug_profile(sess.handle, 3, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
st = buf[: NPH * 4 * G].reshape(NPH, 4, G).astype(np.int64)
tail = buf[NPH * 4 * G:].view(np.uint32)
retr = tail[:NPH]; smid = tail[1024:1024 + G]
arr = st[:, 3, :]
print("smid of CTA 0..147:", smid.tolist())
t_last = arr.max(axis=1); t_med = np.median(arr, axis=1); t_first = arr.min(axis=1)
print("frame us", (t_last[-1] - t_last[0]) / 1e3, " retries", int(retr.sum()))
print("mean (last - median) ns", float((t_last - t_med).mean()), " mean (last - first)", float((t_last - t_first).mean()))
late = (arr - t_med[:, None]).mean(axis=0)
order = np.argsort(late)
print("per-CTA mean lateness at arrive vs median (ns), sorted:")
print([(int(i), int(smid[i]), int(late[i])) for i in order])
kinds = {}
names = ["proj"] + ["qkv", "attn", "o", "gateup", "down"] * 5 + ["head"]
# segment durations per CTA for CP phases of passes 1..14, by phase name: wait-pass -> x ready -> mma done -> arrive, and arrive(prev) -> wait pass
slow = order[-12:]; fast = order[len(order)//2 - 6: len(order)//2 + 6]
print("slow CTAs", slow.tolist(), "median CTAs", fast.tolist())
for nm in ["proj", "qkv", "o", "gateup", "down", "head"]:
    idx = [1 + 27 * g + k for g in range(1, 15) for k in range(27) if names[k] == nm]
    seg = {}
    for label, grp in (("slow", slow), ("med", fast)):
        a0 = st[idx][:, 0, :][:, grp]; a1 = st[idx][:, 1, :][:, grp]; a2 = st[idx][:, 2, :][:, grp]; a3 = st[idx][:, 3, :][:, grp]
        prev = st[[i - 1 for i in idx]][:, 3, :].max(axis=1)[:, None]     # previous phase's last arrive
        seg[label] = (float((a0 - prev).mean()), float((a1 - a0).mean()), float((a2 - a1).mean()), float((a3 - a2).mean()))
    print(f"{nm:7s} [lastarrive->waitpass, x, mma, tail]  slow {[int(v) for v in seg['slow']]}  median {[int(v) for v in seg['med']]}")
