"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier,
HMMA = mma.sync, LDSM = ldmatrix, LDGSTS = cp.async.  Runs without a GPU:  python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qwen3_tts_rs_b200", "libq3tts_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|HMMA|LDSM|LDGSTS)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts, total = None, collections.OrderedDict(), collections.Counter()
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            total[fn] += 1
            for k in PAT.findall(ln):
                counts[fn][k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (sm_100a); kernels without any of them omitted")
    for (f, c), name in zip(counts.items(), demangle):
        if c:
            short = re.sub(r"\(.*", "", name)
            print(f"{short:58s} {total[f]:6d} instr  " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    sys.exit(main())
