"""Stand-alone operators at sizes where HBM bandwidth is visible (for ncu): the fused residual + RMSNorm
custom op (q3_fused_residual_rmsnorm, device pointers) on [65536, 2048] bf16 and [65536, 1024] bf16, and the
per-op sampler (q3_sample) on 256 rows of 3072 logits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from qwen3_tts_rs_b200 import api, lib as L

lib = L.load()
torch.manual_seed(0)
for cols in (2048, 1024):
    rows = 65536
    x = torch.randn(rows, cols, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(rows, cols, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(cols, device="cuda", dtype=torch.bfloat16)
    normed, total = torch.empty_like(x), torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(4):
        if it == 1:
            ev[0].record()
        L.check(lib.q3_fused_residual_rmsnorm(C.c_void_p(x.data_ptr()), C.c_void_p(r.data_ptr()), C.c_void_p(w.data_ptr()),
                                              C.c_void_p(normed.data_ptr()), C.c_void_p(total.data_ptr()), rows, cols,
                                              1e-6, L.Q3_BF16, C.c_void_p(st)))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    gb = 4 * rows * cols * 2 / 1e9
    print(f"fused_residual_rmsnorm bf16 [{rows},{cols}]: {ms*1e3:.1f} us, {gb/ms*1e3:.0f} GB/s algorithmic (2 reads + 2 writes)")

rng = np.random.default_rng(0)
B, V = 256, 3072
logits = rng.standard_normal((B, V)).astype(np.float32) * 3
states = np.arange(1, B + 1, dtype=np.uint64)
seen = np.zeros((B, V), dtype=np.uint8)
tok = api.sample(logits, api.SynthesisOptions(), states, seen, 5)
print("sample tokens", tok[:8])
