"""TEST INFRASTRUCTURE ONLY.  Build recipe for oracle/_ref/ and oracle/c/.

oracle/_ref/libref_fused_rmsnorm.so  -- the REFERENCE's own CUDA kernel
    (/root/reference/kernels/fused_residual_rmsnorm.cu, the only native kernel in the reference),
    compiled from where it lies for sm_100a plus a 40-line host launcher (oracle/ref_launcher.cu).
    Flags: --use_fast_math, because the PTX the reference ships and JIT-loads at run time
    (kernels/fused_residual_rmsnorm.ptx: fma.rn.ftz, div.approx.ftz, rsqrt.approx.ftz) was built that
    way.  Only possible where /root/reference exists (this container); the GPU box uses the prebuilt
    .so, which is git-ignored but travels with the gpurun snapshot.  No reference source is copied.
oracle/c/libq3oracle_c.so  -- the plain-C restatement of the integer / byte parts of the path
    (PCG RNG, suppression rule, codes_to_tensor, PCM16 conversion), built with gcc.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_KERNEL = "/root/reference/kernels/fused_residual_rmsnorm.cu"
REF_DIR = os.path.join(HERE, "_ref")
REF_LIB = os.path.join(REF_DIR, "libref_fused_rmsnorm.so")
REF_CUBIN = os.path.join(REF_DIR, "fused_residual_rmsnorm_sm100a.cubin")
C_LIB = os.path.join(HERE, "c", "libq3oracle_c.so")


def build_ref(force: bool = False):
    """Returns the path of the reference-kernel library, or None when it cannot be (re)built here and
    no prebuilt copy exists."""
    if not os.path.exists(REF_KERNEL):
        return REF_LIB if (os.path.exists(REF_LIB) and os.path.exists(REF_CUBIN)) else None
    os.makedirs(REF_DIR, exist_ok=True)
    src = os.path.join(HERE, "ref_launcher.cu")
    if not force and os.path.exists(REF_LIB) and os.path.getmtime(REF_LIB) > max(os.path.getmtime(src), os.path.getmtime(REF_KERNEL)):
        return REF_LIB
    # 1. the reference kernel, unmodified, device code only (it guards its bf16/f16 kernels with
    #    __CUDA_ARCH__, so it is meant to be compiled to device code and loaded as a module)
    cmd = ["/usr/local/cuda/bin/nvcc", "-arch=sm_100a", "--use_fast_math", "-O3", "-cubin", "-o", REF_CUBIN, REF_KERNEL]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference kernel failed:\n" + r.stderr[-3000:])
    # 2. the launcher
    cmd = ["/usr/local/cuda/bin/nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", REF_LIB, src, "-lcudart_static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference launcher failed:\n" + r.stderr[-3000:])
    return REF_LIB


def build_c(force: bool = False):
    src = os.path.join(HERE, "c", "q3_oracle.c")
    if not force and os.path.exists(C_LIB) and os.path.getmtime(C_LIB) > os.path.getmtime(src):
        return C_LIB
    r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", C_LIB, src, "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building oracle/c failed:\n" + r.stderr[-3000:])
    return C_LIB


if __name__ == "__main__":
    print(build_c())
    print(build_ref())
