#!/usr/bin/env python
"""Overlap statistics of an ICB_TRACE file (lines: kernel stream start_ms end_ms).

    ICB_TRACE=gpurun_out/trace.txt python tools/bench_workers.py 1000000 16 1 trace ; python tools/trace_stats.py gpurun_out/trace.txt
"""
import collections
import sys

import numpy as np

rows = [l.split() for l in open(sys.argv[1]) if l.strip()]
names = sorted(set(r[0] for r in rows))
a = np.array([float(r[2]) for r in rows])
b = np.array([float(r[3]) for r in rows])
nm = np.array([r[0] for r in rows])
t0, t1 = a.min(), b.max()
print(f"{len(rows)} launches over {t1 - t0:.1f} ms, {len(set(r[1] for r in rows))} streams")
for n in names:
    m = nm == n
    print(f"  {n:22s} launches {m.sum():7d}  sum {np.sum(b[m] - a[m]):10.1f} ms  avg {np.mean(b[m] - a[m]) * 1e3:8.1f} us")


def coverage(mask, label):
    ev = sorted([(x, 1) for x in a[mask]] + [(x, -1) for x in b[mask]])
    depth, last, busy, area = 0, t0, 0.0, 0.0
    hist = collections.Counter()
    for t, d in ev:
        if depth > 0:
            busy += t - last
        area += depth * (t - last)
        hist[min(depth, 40)] += t - last
        depth += d
        last = t
    print(f"{label}: some launch in flight {busy:.1f} ms ({100 * busy / (t1 - t0):.1f} % of the span), mean depth {area / (t1 - t0):.2f}")
    tot = sum(hist.values())
    print("   time share by launches in flight:", " ".join(f"{k}:{100 * v / tot:.0f}%" for k, v in sorted(hist.items()) if v / tot > 0.005))


coverage(np.ones(len(rows), bool), "all kernels")
for n in names:
    coverage(nm == n, n)
