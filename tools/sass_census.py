#!/usr/bin/env python
"""Dev tool: instruction census of one kernel per SOURCE LINE, split by issue pipe (no GPU needed).

    python tools/sass_census.py isonclust2_b200/csrc/sketch.cu icb_k_sketch_warpILi5ELi2688ELi4ELi11E [min_count]

Compiles the file to a cubin for sm_100a with -lineinfo, disassembles it with `nvdisasm --print-line-info` and attributes
every SASS instruction of the kernel whose mangled name contains the pattern to the innermost line of THIS source file in its
inline chain.  ALU = the integer / logic pipe (LOP3, SHF, ISETP, VIADD, IADD3, SEL, VIMNMX*, LEA, PRMT, ...: 16 lanes per
clock, tools/microbench/pipe_rates.cu), IMAD = the FMA pipe.  Hot loops that are fully unrolled (the sketch and alignment
kernels') make the static count times the trip count a usable model of the executed mix; ncu's
sm__inst_executed_pipe_* is the check (profiles/r02_summary.md, "Sketch kernel, last change of the round").
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALU = re.compile(r"^(LOP3|SHF|ISETP|VIADD|IADD3|SEL|VIMNMX3|VIMNMX|LEA|PRMT|VIADDMNMX|PLOP3|POPC|FLO|R2P|P2R|BREV|SGXT|BMSK)$")


def main():
    src, pat = sys.argv[1], sys.argv[2]
    thr = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    base = os.path.basename(src)
    with tempfile.TemporaryDirectory() as d:
        cubin = os.path.join(d, "k.cubin")
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "isonclust2_b200", "csrc"), "-cubin", "-x", "cu",
                               src, "-o", cubin])
        txt = subprocess.check_output(["nvdisasm", "--print-line-info", cubin], text=True).split("\n")
    starts = [i for i, l in enumerate(txt) if l.startswith("//------") and ".text." in l and pat in l]
    if not starts:
        sys.exit(f"no kernel matching {pat}")
    start = starts[0]
    end = next(i for i, l in enumerate(txt) if l.startswith("//------") and i > start)
    cur = None
    cnt, alu, fma = collections.Counter(), collections.Counter(), collections.Counter()
    for ln in txt[start:end]:
        if "//## File" in ln:
            ms = [int(x) for x in re.findall(re.escape(base) + r'", line (\d+)', ln)]
            cur = ms[-1] if ms else cur
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m and cur is not None:
            op = m.group(2).split(".")[0]
            cnt[cur] += 1
            alu[cur] += bool(ALU.match(op))
            fma[cur] += op == "IMAD"
    lines = open(src).read().split("\n")
    print("TOTAL", sum(cnt.values()), "ALU", sum(alu.values()), "IMAD", sum(fma.values()))
    for k in sorted(cnt):
        if cnt[k] >= thr:
            print(f"{k:5d} n={cnt[k]:4d} alu={alu[k]:4d} imad={fma[k]:3d} | {lines[k - 1].strip()[:100]}")


if __name__ == "__main__":
    main()
