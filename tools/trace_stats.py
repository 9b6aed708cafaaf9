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

# where the chain of ONE stream spends its time: kernels by name and the gaps between consecutive launches
streams = collections.Counter(r[1] for r in rows)
print("\nper-stream chains (the three streams with the most launches):")
for sid, cnt in streams.most_common(3):
    m = np.array([r[1] == sid for r in rows])
    order = np.argsort(a[m])
    sa, sb, sn = a[m][order], b[m][order], nm[m][order]
    span = sb.max() - sa.min()
    gaps = np.maximum(0.0, sa[1:] - np.maximum.accumulate(sb)[:-1])
    print(f" stream {sid}: {cnt} launches over {span:.1f} ms; in kernels {np.sum(sb - sa):.1f} ms; gaps {gaps.sum():.1f} ms "
          f"(gaps > 50 us: {np.sum(gaps > 0.05)} totalling {gaps[gaps > 0.05].sum():.1f} ms)")
    for n in names:
        k = sn == n
        if k.any():
            print(f"     {n:22s} {k.sum():6d} launches {np.sum(sb[k] - sa[k]):9.1f} ms  avg {np.mean(sb[k] - sa[k]) * 1e3:8.1f} us  p90 {np.percentile(sb[k] - sa[k], 90) * 1e3:8.1f} us")
