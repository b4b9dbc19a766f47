"""Race hunt, step 3: phase n = 10 (layer-1 o_proj of pass 0): which inputs / outputs differ between repetitions?"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W
spec = S.SPECS["1.7b"]; B = 8; n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=64), list(range(B)), max_seq=128)
lib = L.load()
fn = lib.q3_debug_cp_prefix
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 5
fn.restype = C.c_int
g = torch.Generator().manual_seed(5)
hid = (torch.randn(B, spec.hidden, generator=g) * 0.7).to(torch.bfloat16).contiguous()
toks = np.asarray([100 + 37 * i for i in range(B)], dtype=np.uint32)
Hm, Im = 2048, 6144; nhm = 32 * 128; qdm = 16 * 128
bufs = {k: np.zeros(m, dtype=np.uint32) for k, m in (("x", 16 * Hm), ("qkv", 16 * nhm), ("attn", 16 * qdm), ("h1", 16 * Hm * 2), ("act", 16 * Im))}
def run():
    L.check(fn(sess.handle, hid.data_ptr(), toks.ctypes.data, n, *[bufs[k].ctypes.data for k in ("x", "qkv", "attn", "h1", "act")]))
    return {k: v.reshape(-1, 2).copy() for k, v in bufs.items()}
ref = run()
C_ = spec.cp_hidden
for r in range(12):
    cur = run()
    line = []
    for k in ("x", "attn", "h1"):
        same = np.array_equal(cur[k][:, 0], ref[k][:, 0])
        line.append(f"{k}:{'same' if same else 'DIFF'}")
    h_c, h_r = cur["h1"][:16 * C_, 0].view(np.float32).reshape(16, C_), ref["h1"][:16 * C_, 0].view(np.float32).reshape(16, C_)
    tags = cur["h1"][:16 * C_, 1].reshape(16, C_)
    d = np.abs(h_c - h_r)
    cols = np.nonzero(d.max(0) > 0)[0]
    toks_d = np.nonzero(d.max(1) > 0)[0]
    line.append(f"h1 tokens differing {toks_d.tolist()} cols differing {len(cols)} (CTA row blocks {sorted(set((cols // 8).tolist()))[:20]}...) max|d| {d.max():.4f} tags uniq {np.unique(tags).tolist()[:4]}")
    print(f"rep {r}: " + " ".join(line), flush=True)
