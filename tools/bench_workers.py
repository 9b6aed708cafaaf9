#!/usr/bin/env python
"""Sweep of `n_workers` (batches in flight per GPU) for the device-resident pipeline step of bench.py.

    python tools/bench_workers.py [reads] [workers,workers,...] [passes]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402
import torch  # noqa: E402

from isonclust2_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sweep = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,32,48").split(",")]
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
profile = len(sys.argv) > 4 and sys.argv[4] == "profile"
trace = len(sys.argv) > 4 and sys.argv[4] == "trace"  # per-launch event timing on (with ICB_TRACE=<file>: launch intervals are logged)  # bracket the timed passes with cudaProfilerStart/Stop (ncu --profile-from-start off)
seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
ctx = api.Context(0)
rs = ctx.upload(seq, qual, offs)
ctx.sync()
api.run_pipeline(ctx, None, None, None, rs=rs, n_workers=sweep[0])
if profile:
    torch.cuda.profiler.start()
if trace:
    ctx.timing(True)
for nw in sweep:
    ts = []
    for _ in range(passes):
        ctx.sync()
        t0 = time.perf_counter()
        res = api.run_pipeline(ctx, None, None, None, rs=rs, n_workers=nw)
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    st = res["stats"]
    print(f"workers={nw:3d}  {min(ts):7.3f} s  {n / min(ts):9.0f} reads/s  rounds={st['rounds']} pairs={st['pairs_aligned']} "
          f"ms: query={st['ms_query']:.0f} full={st['ms_full']:.0f} align={st['ms_align']:.0f} insert={st['ms_insert']:.0f} host={st['ms_host']:.0f} cpu={st['ms_cpu']:.0f} order={st['ms_order']:.0f}/{st['order_calls']} decisions={st['decisions']}",
          flush=True)
if profile:
    torch.cuda.profiler.stop()
if trace:
    ctx.timing(False)
rs.close()
ctx.close()
