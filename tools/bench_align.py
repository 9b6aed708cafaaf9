#!/usr/bin/env python
"""Dev tool: time icb_align_pairs on realistic (read, representative) pairs.  usage: bench_align.py [n_reads] [pairs_per_call] [reps]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from isonclust2_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
per = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
prof = len(sys.argv) > 4 and sys.argv[4] == "prof"  # bracket the saturated call with cudaProfilerStart/Stop (ncu --profile-from-start off)
k, w = 11, 15
seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
ctx = api.Context(0)
rs = ctx.upload(seq, qual, offs)
res = api.run_pipeline(ctx, None, None, None, rs=rs, k=k, w=w, mode="sahlin", batch_kb=0, batch_max_seq=0)
order, cl, st = res["order"], res["cluster"], res["strand"]
founder = {}
rows = []
for s in range(n):
    c = int(cl[s])
    if c < 0:
        continue
    if c not in founder:
        founder[c] = int(order[s])
        continue
    rows.append((int(order[s]), founder[c], int(st[s]) or 1, 3, 8))
pairs = np.array(rows, dtype=api.PAIR_DT)
print("pairs available", len(pairs))
lens = np.diff(offs)
for per_call in (per, len(pairs)):
    sub = pairs[:per_call]
    cells = float((lens[sub["read"]] * lens[sub["rep"]]).sum())
    rs.align_pairs(sub, k)
    if prof and per_call == len(pairs):
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaProfilerStart()
        rs.align_pairs(sub, k)
        rt.cudaProfilerStop()
    ctx.timing(True); ctx.timing_reset()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = rs.align_pairs(sub, k)
    dt = (time.perf_counter() - t0) / reps
    kd, kt = ctx.kernel_time("align_dp"), ctx.kernel_time("align_trace")
    ctx.timing(False)
    # a call may split into several launches (workspace limit): DP / trace time per CALL = summed launch time / reps
    dp_ms, tr_ms = kd[1] / reps, kt[1] / reps
    print(f"pairs/call {len(sub)}: wall {dt*1e3:.3f} ms  dp {dp_ms:.3f} ms in {kd[0] // reps} launch(es)  trace {tr_ms:.3f} ms  "
          f"GCUPS(dp) {cells/(dp_ms*1e-3)/1e9:.1f}  GCUPS(wall) {cells/dt/1e9:.1f}  mean aligned {out['aligned'].mean():.1f}")
