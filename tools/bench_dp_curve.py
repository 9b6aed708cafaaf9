#!/usr/bin/env python
"""Dev tool: throughput of the packed alignment kernel against the number of CTAs per SM, on pairs of (nearly) one length.
usage: bench_dp_curve.py [lo] [hi]   (read lengths taken from [lo, hi])"""
import sys
import numpy as np
sys.path.insert(0, '.')
from isonclust2_b200 import api, synth

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 1300
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 1400
seq, qual, offs = synth.make_reads(6000, seed=100, err_set=(0.08,))
lens = np.diff(offs)
pick = np.nonzero((lens >= lo) & (lens <= hi))[0]
print("reads in range", len(pick), "mean length", lens[pick].mean())
ctx = api.Context(0)
rs = ctx.upload(seq, qual, offs)
api.run_pipeline(ctx, None, None, None, rs=rs, k=11, w=15, mode="fast", batch_kb=0, batch_max_seq=0)  # scores / sketches exist
rng = np.random.default_rng(1)
for per_call in (74, 148, 296, 444, 592, 1184, 2368, 4736):
    a = rng.choice(pick, per_call)
    b = rng.choice(pick, per_call)
    pairs = np.zeros(per_call, dtype=api.PAIR_DT)
    pairs["read"], pairs["rep"], pairs["strand"], pairs["gap_open"], pairs["limit"] = a, b, 1, 3, 8
    cells = float((lens[a] * lens[b]).sum())
    rs.align_pairs(pairs, 11)
    ctx.timing(True); ctx.timing_reset()
    reps = 5
    for _ in range(reps):
        rs.align_pairs(pairs, 11)
    kd, kt = ctx.kernel_time("align_dp"), ctx.kernel_time("align_trace")
    ctx.timing(False)
    print(f"pairs {per_call:5d} ({per_call / 148:.1f} per SM): dp {kd[1] / reps:.3f} ms in {kd[0] // reps} launch(es), {cells / (kd[1] / reps * 1e-3) / 1e9:.0f} GCUPS; "
          f"trace {kt[1] / reps:.3f} ms")
