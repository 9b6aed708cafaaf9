#!/usr/bin/env python
"""Dev tool: how many CTAs of the packed alignment kernel each SM holds over one step of the bench workload.

usage: residency.py [n_reads] [out.json]   (uses the library's residency log, include/isonclust2_b200.h "Diagnostics")
"""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from isonclust2_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
out = sys.argv[2] if len(sys.argv) > 2 else None
seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
ctx = api.Context(0)
rs = ctx.upload(seq, qual, offs)
pipe = dict(k=11, w=15, mode="sahlin", batch_kb=50000, batch_max_seq=30000)
for _ in range(2):
    res = api.run_pipeline(ctx, None, None, None, rs=rs, **pipe)
ctx.sync()
cap = 8 << 20
ctx.check(ctx.L.icb_debug_residency_start(ctx.h, cap))
res = api.run_pipeline(ctx, None, None, None, rs=rs, **pipe)
ctx.sync()
buf = np.zeros((cap, 4), dtype=np.uint32)
got = C.c_int64()
ctx.check(ctx.L.icb_debug_residency_dump(ctx.h, buf.ctypes.data, cap, C.byref(got)))
rec = buf[: got.value]
kind = rec[:, 0] & 0xF
sm = (rec[:, 0] >> 4) & 0xFF
t0 = rec[:, 1].astype(np.int64)
t1 = rec[:, 2].astype(np.int64)
base = t0.min()
t0 -= base
t1 -= base
span = int(t1.max())
rep = {"records": int(got.value), "step_us": span * 1.024, "sms": int(sm.max()) + 1}
for kd, name, per_sm_max in ((1, "dp", 4), (2, "trace_warps", 64)):
    sel = kind == kd
    if not sel.any():
        continue
    a0, a1, s_ = t0[sel], t1[sel], sm[sel]
    dur = (a1 - a0) * 1.024
    cells = (rec[sel, 3] & 0xFFFF).astype(np.float64) * (rec[sel, 3] >> 16).astype(np.float64)
    rep[name] = {"n": int(sel.sum()), "mean_us": float(dur.mean()), "p50_us": float(np.median(dur)), "p90_us": float(np.percentile(dur, 90)),
                 "resident_seconds_total": float(dur.sum() * 1e-6), "cells": float(cells.sum())}
    # residency histogram: for every SM, time spent with c units resident
    nsm = int(sm.max()) + 1
    hist = np.zeros(per_sm_max + 2)
    timeline = np.zeros(50)  # mean resident units per SM in 50 slices of the step
    for s in range(nsm):
        m = s_ == s
        ev = np.concatenate([np.stack([a0[m], np.ones(m.sum(), dtype=np.int64)], 1), np.stack([a1[m], -np.ones(m.sum(), dtype=np.int64)], 1)])
        ev = ev[np.lexsort((ev[:, 1], ev[:, 0]))]
        cur, last = 0, 0
        for t, dlt in ev:
            if t > last:
                hist[min(cur, per_sm_max + 1)] += t - last
                lo, hi = last * 50 // (span + 1), t * 50 // (span + 1)
                for b in range(lo, hi + 1):
                    b0, b1 = max(last, b * (span + 1) // 50), min(t, (b + 1) * (span + 1) // 50)
                    if b1 > b0:
                        timeline[b] += cur * (b1 - b0)
                last = t
            cur += dlt
        hist[0] += span - last
    rep[name]["time_share_by_resident_count"] = [round(float(x), 4) for x in hist / hist.sum()]
    rep[name]["mean_resident_per_sm_over_50_slices"] = [round(float(x), 2) for x in timeline / (nsm * (span + 1) / 50)]
    if kd == 1:
        # cell rate of a CTA by how long it lived: cells / duration
        rep[name]["gcups_per_cta_mean"] = float((cells / (dur * 1e-6)).mean() / 1e9)
print(json.dumps(rep))
if out:
    json.dump(rep, open(out, "w"), indent=1)
