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


# Two libraries from the same sources:
#   libq3tts_b200.so      the product: the dataflow kernel (mega2.cuh, the default) and the two TMA-fed generations (mega4.cuh,
#                         mega5.cuh, opt-in), profiling hooks compiled out (code that never runs still costs instruction fetch
#                         in a run-once-per-phase kernel)
#   libq3tts_b200_dev.so  + the historical generations (Q3_MEGA=1 / 3) and the profiling hooks (tools/profile_*.py, and the
#                         tests that keep the old generations against the oracle); selected with Q3TTS_LIB=dev
VARIANTS = {"": [], "_dev": ["-DQ3_ALL_GENERATIONS=1", "-DQ3_PROF=1"]}


def lib_path(variant: str = "") -> str:
    return os.path.join(HERE, f"libq3tts_b200{variant}.so")


def build(force: bool = False, verbose: bool = False, variants=None) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "q3tts.h"))
    variants = list(VARIANTS) if variants is None else variants
    jobs, objs = [], {v: [] for v in variants}
    for v in variants:
        for src in SOURCES:
            sp = os.path.join(CSRC, src)
            obj = os.path.join(CSRC, src.replace(".cu", f"{v}.o"))
            objs[v].append(obj)
            if force or _stale(obj, [sp] + headers):
                jobs.append((sp, obj, VARIANTS[v]))

    def compile_one(job):
        sp, obj, defs = job
        cmd = [_nvcc()] + NVCC_FLAGS + defs + ["-c", sp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(CSRC, os.path.basename(obj) + ".ptxas.log")
        with open(log, "w") as f:      # register / spill report (git-ignored build artefact)
            f.write("".join(ln for ln in r.stderr.splitlines(True) if "Compile time" not in ln))
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{r.stderr[-6000:]}")
        if verbose:
            print(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(compile_one, jobs))
    for v in variants:
        lib = lib_path(v)
        if force or jobs or _stale(lib, objs[v]):
            cmd = [_nvcc(), "-shared", "-o", lib] + objs[v] + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
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
