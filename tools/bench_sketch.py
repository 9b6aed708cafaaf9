#!/usr/bin/env python
"""Dev tool: time the stage-1 kernels alone.  usage: bench_sketch.py [n_reads] [reps]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from isonclust2_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
k, w = 11, 15
seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
ctx = api.Context(0)
rs = ctx.upload(seq, qual, offs)
score, err = rs.score(k)
skip = api.quality_filtered(err, 7.0)
rs.sketch_async(k, w, skip=skip); ctx.sync()
ctx.timing(True); ctx.timing_reset()
for _ in range(reps):
    rs.score_async(k)
    rs.sketch_async(k, w, skip=skip)
ctx.sync()
nf, nr = rs.sketch_sizes()
sk = rs.sketch_download(hpc=False)
hl = int(sk["hpc_len"].sum())
nb = int(offs[-1])
alg = 2 * nb + 2 * hl + 12 * (nf + nr) + 24 * n
for name in ("sketch_reads", "score_reads", "hpc_error_rate", "sketch_reads_cta", "sketch_reads_long"):
    c, ms = ctx.kernel_time(name)
    if c:
        print(f"{name}: {ms/c:.3f} ms/launch" + (f"  -> {alg/(ms/c*1e-3)/1e9:.0f} GB/s algorithmic ({alg/1e6:.0f} MB)" if name == "sketch_reads" else ""))
print("reads", n, "bases", nb, "hpc", hl, "mins", nf, nr)
