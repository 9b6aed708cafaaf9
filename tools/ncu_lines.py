#!/usr/bin/env python
"""Aggregate an ncu report's executed instructions / stall samples per CUDA source line.

usage: tools/ncu_lines.py report.ncu-rep [top_n]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg, src, cur = collections.OrderedDict(), {}, None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= iI:
        continue
    if r[0] != "":
        try:
            cur = int(r[0])
            src[cur] = r[1]
        except ValueError:
            continue
    try:
        v, sm = int(r[iI]), int(r[iS] or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0])
    a[0] += v
    a[1] += sm
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("total warp instructions", tot, "samples", tots)
for ln, (v, sm) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{ln:5d} inst {100 * v / tot:5.1f}%  samples {100 * sm / tots:5.1f}%  {src.get(ln, '').strip()[:110]}")
