#!/usr/bin/env python
"""Dev tool: where the rounds of a batch spend their time WITH every batch in flight and the round graphs on (no event timing):
the kernels of the device loop mark their start in the library's residency log (csrc/residency.cuh); consecutive marks of one
batch give the time from the start of one kernel to the start of the next.

usage: round_timeline.py [n_reads] [out.json]
"""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from isonclust2_b200 import api, synth  # noqa: E402

KIND = {3: "commit", 4: "decide", 5: "order", 6: "order_large", 7: "gather", 8: "query_full", 9: "align_dp", 10: "align_trace", 11: "align_dp32",
        12: "query_pool", 13: "finalize", 14: "init"}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
out = sys.argv[2] if len(sys.argv) > 2 else None
seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
ctx = api.Context(0)
rs = ctx.upload(seq, qual, offs)
pipe = dict(k=11, w=15, mode="sahlin", batch_kb=50000, batch_max_seq=30000)
for _ in range(2):
    api.run_pipeline(ctx, None, None, None, rs=rs, **pipe)
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
tag = rec[:, 0] >> 12
t0 = rec[:, 1].astype(np.int64)
base = t0.min()
marks = kind >= 3
rep = {"records": int(got.value), "marks": int(marks.sum()), "step_ms": float((rec[:, 2].astype(np.int64).max() - base) * 1.024e-3), "batches": {}}
tags, counts = np.unique(tag[marks], return_counts=True)
order = np.argsort(-counts)
summary = {}
for ti in order:
    tg = tags[ti]
    m = marks & (tag == tg)
    ts = (t0[m] - base) * 1.024e-3  # ms
    ks = kind[m]
    o = np.argsort(ts, kind="stable")
    ts, ks = ts[o], ks[o]
    if len(ts) < 50:
        continue
    dt = np.diff(ts)
    stage = {}
    for kd, name in KIND.items():
        sel = ks[:-1] == kd
        if sel.any():
            stage[name] = {"n": int(sel.sum()), "total_ms": float(dt[sel].sum()), "mean_us": float(dt[sel].mean() * 1e3), "p90_us": float(np.percentile(dt[sel], 90) * 1e3)}
    rounds = int((ks == 9).sum())
    summary[int(tg)] = {"marks": int(len(ts)), "rounds": rounds, "span_ms": float(ts[-1] - ts[0]), "first_ms": float(ts[0]), "last_ms": float(ts[-1]), "stages": stage}
rep["batches"] = summary
# aggregate over batches
agg = {}
for b in summary.values():
    for name, s in b["stages"].items():
        a = agg.setdefault(name, [0, 0.0])
        a[0] += s["n"]
        a[1] += s["total_ms"]
tot = sum(a[1] for a in agg.values())
print(f"step {rep['step_ms']:.0f} ms, {len(summary)} batches, {sum(b['rounds'] for b in summary.values())} rounds")
print("time from the start of a kernel to the start of the next one of the same batch, summed over all batches:")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name:12s} {a[0]:7d} x {a[1] / a[0] * 1e3:8.1f} us = {a[1]:9.1f} ms  {100 * a[1] / tot:5.1f} %")
longest = max(summary.values(), key=lambda b: b["span_ms"])
print(f"longest batch: {longest['rounds']} rounds over {longest['span_ms']:.0f} ms ({longest['span_ms'] / max(1, longest['rounds']):.2f} ms per round)")
for name, s in sorted(longest["stages"].items(), key=lambda kv: -kv[1]["total_ms"]):
    print(f"  {name:12s} {s['n']:6d} x {s['mean_us']:8.1f} us (p90 {s['p90_us']:8.1f}) = {s['total_ms']:8.1f} ms")
if out:
    json.dump(rep, open(out, "w"), indent=1)
