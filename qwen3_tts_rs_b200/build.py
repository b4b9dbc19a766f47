"""Build libq3tts_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libq3tts_b200.so")
SOURCES = ["q3tts.cu", "vocoder.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "q3tts.h"))
    objs = []
    jobs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            jobs.append((sp, obj))

    def compile_one(job):
        sp, obj = job
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", sp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(CSRC, os.path.basename(obj) + ".ptxas.log")
        with open(log, "w") as f:      # register / spill report, kept in the tree; compile times dropped so it is stable
            f.write("".join(ln for ln in r.stderr.splitlines(True) if "Compile time" not in ln))
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{r.stderr[-6000:]}")
        if verbose:
            print(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(compile_one, jobs))
    if force or jobs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    return LIB


def build_host_example() -> str:
    """Compile the C++ host mirror's example / test driver (include/q3tts.hpp over the C ABI) with g++ against the
    freshly built library: the "does it build" check of the compiled host side."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tests", "cpp", "host_mirror_check.cpp")
    out = os.path.join(root, "tests", "cpp", "host_mirror_check")
    deps = [src, os.path.join(root, "include", "q3tts.hpp"), os.path.join(root, "include", "q3tts.h"), LIB]
    if _stale(out, deps):
        cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(root, "include"), src, "-o", out,
               "-L", HERE, "-lq3tts_b200", f"-Wl,-rpath,{HERE}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {src}:\n{r.stderr[-4000:]}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
