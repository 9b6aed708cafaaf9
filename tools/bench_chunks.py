#!/usr/bin/env python
"""Sweep of the chunk size (reads scored against one index snapshot) for the device-resident pipeline step of bench.py.

    python tools/bench_chunks.py [reads] [chunk,chunk,...] [passes]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from isonclust2_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sweep = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "128,256,512,1024").split(",")]
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
ctx = api.Context(0)
rs = ctx.upload(seq, qual, offs)
ctx.sync()
ref = None
for cs in sweep:
    ts = []
    for _ in range(passes):
        ctx.sync()
        t0 = time.perf_counter()
        res = api.run_pipeline(ctx, None, None, None, rs=rs, chunk_size=cs)
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    st = res["stats"]
    same = True if ref is None else bool((ref["cluster"] == res["cluster"]).all() and (ref["strand"] == res["strand"]).all())
    ref = ref or res
    nb = len(res["batches"])
    print(f"chunk={cs:5d}  {min(ts):7.3f} s  {n / min(ts):9.0f} reads/s  same_result={same} chunks={st['chunks']} rounds={st['rounds']} "
          f"pairs={st['pairs_aligned']} per-batch ms: query={st['ms_query'] / nb:.0f} full={st['ms_full'] / nb:.0f} align={st['ms_align'] / nb:.0f} "
          f"insert={st['ms_insert'] / nb:.0f} host={st['ms_host'] / nb:.0f} order={st['ms_order'] / nb:.0f} cpu={st['ms_cpu'] / nb:.0f} "
          f"decisions={st["decisions"]} hyp={st["hyp_decisions"]} order_calls={st['order_calls']} slowest_batch={st['ms_batch_max']:.0f} cluster_wall={st['ms_wall']:.0f} "
          f"phases={ {k: round(v) for k, v in res['phases_ms'].items()} }", flush=True)
rs.close()
ctx.close()
