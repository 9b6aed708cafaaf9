#!/usr/bin/env python
"""bench.py — isonclust2_b200 hot-path benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl ours|reference]

One "step" = one pass of the clustering hot path over one batch set of synthetic ONT-like cDNA reads
(BASELINE.json configs[1]: ~1.3 kb mean, 8 % error, k=11 w=15): quality scores -> stable sort -> HPC minimizer
sketches -> batches (main.cpp:149-199 rule) -> greedy clustering of every batch in `sahlin` mode (device index
queries + batched alignment verification).  metric = reads/s clustered.
  value : reads already resident in HBM when the timed region starts (CUDA events on the library's stream)
  e2e   : through the public C ABI with pinned HOST buffers; H2D of the reads and D2H of the per-read results are
          inside the timed region (wall clock bracketed by device synchronisation)
Multi-GPU: reads are sharded across ranks, one process per GPU, no data-path collective -> weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, W_, MIN_QUAL, MODE = 11, 15, 7.0, "sahlin"
BATCH_KB, BATCH_MAX_SEQ = 50000, 30000  # CmdArgs defaults (args.h:14-15)
METRIC = "reads/sec clustered (sahlin)"
KERNELS = ("sketch_reads", "score_reads", "hpc_error_rate", "index_query", "index_query_full", "index_insert", "align_dp",
           "align_trace", "align_dp32", "classify_errors", "index_rehash", "sketch_reads_long", "gather_minimizers")


def clocks_sampler(stop, out, device):
    """nvidia-smi clock/throttle samples during the timed region (B200_PROFILING.md recipe)."""
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    num = lambda s: s.replace(".", "", 1).isdigit()
    sm = sorted(float(s[0]) for s in samples if num(s[0]))
    mx = [float(s[1]) for s in samples if num(s[1])]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from isonclust2_b200 import api, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_reads = args.reads  # per GPU (weak scaling)
    pin = lambda n: torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    seq, qual, offs = synth.make_reads(n_reads, seed=100 + rank, err_set=(0.08,), seq_out=pin, qual_out=pin)
    n_bases = int(offs[-1])

    ctx = api.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pipe = dict(k=K, w=W_, mode=MODE, min_qual=MIN_QUAL, batch_kb=BATCH_KB, batch_max_seq=args.batch_max_seq, chunk_size=args.chunk)

    # ---------------- device-resident steps: reads already in HBM
    rs = ctx.upload(seq, qual, offs)
    ctx.sync()
    res = None
    for _ in range(args.warmup):
        res = api.run_pipeline(ctx, None, None, None, rs=rs, **pipe)
    ctx.sync()

    samples, stop = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    th.start()

    ctx.timing(True)
    ctx.timing_reset()
    launches0 = ctx.launch_count()
    times = []
    for _ in range(args.steps):
        l2_flush.add_(1)  # flush L2 between timed iterations (256 MiB > 126 MB)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = api.run_pipeline(ctx, None, None, None, rs=rs, **pipe)
        e1.record(stream)
        ctx.sync()
        barrier()
        times.append(e0.elapsed_time(e1))
    launches = ctx.launch_count() - launches0
    kt = {n: ctx.kernel_time(n) for n in KERNELS}
    ctx.timing(False)
    nf, nr = rs.sketch_sizes()
    sk = rs.sketch_download(hpc=False)
    hpc_total = int(sk["hpc_len"].sum())
    rs.close()

    # ---------------- e2e through the C ABI with host buffers
    for _ in range(1):
        api.run_pipeline(ctx, seq, qual, offs, **pipe)
    e2e_times = []
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        api.run_pipeline(ctx, seq, qual, offs, **pipe)
        ctx.sync()
        barrier()
        e2e_times.append((time.perf_counter() - t0) * 1e3)
    stop.set()
    th.join(timeout=2)

    ms = float(np.mean(times))
    ms_e2e = float(np.mean(e2e_times))
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # roofline of the dominant data-moving kernel (sketch_reads): algorithmic bytes per launch / measured launch time
    steps = max(args.steps, 1)
    alg_bytes = 2 * n_bases + 2 * hpc_total + 12 * (nf + nr) + 24 * n_reads
    n_l, tot_ms = kt["sketch_reads"]
    k_ms = tot_ms / max(n_l, 1)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    st = res["stats"]
    dp_ms = kt["align_dp"][1] / steps
    gcups = st["cells_aligned"] / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else 0.0
    clustered = int((res["cluster"] >= 0).sum())

    line = {
        "metric": METRIC, "value": world * n_reads / (ms * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u32/i32 (+f64 quality chains)", "data": "synthetic",
        "config": {"workload": f"configs[1] read set: {n_reads} synthetic ONT cDNA reads per GPU (~1.3 kb mean, 8% error), k=11 w=15, "
                               f"sort + cluster -x sahlin of every batch (B={BATCH_KB} kb, M={args.batch_max_seq})",
                   "reads_per_gpu": n_reads, "bases_per_gpu": n_bases, "batches": len(res["batches"]),
                   "clusters": int(sum(res["n_clusters"])), "reads_clustered": clustered,
                   "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"reads sharded x{world}, no collective"},
        "e2e": {"value": world * n_reads / (ms_e2e * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * n_bases + 8 * (n_reads + 1) + n_reads, "d2h_bytes_per_step": 16 * n_reads + 24 * n_reads},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "icb_k_sketch (sketch_reads)", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel_ms": k_ms},
        "alignment": {"gcups": gcups, "cells_per_step": st["cells_aligned"], "pairs_per_step": st["pairs_aligned"],
                      "aln_invoked": st["aln_invoked"], "dp_ms_per_step": dp_ms},
        "cluster_stats": st,
        "kernels_ms_per_step": {n: v[1] / steps for n, v in kt.items() if v[0]},
        "kernel_launches_per_step": {n: v[0] / steps for n, v in kt.items() if v[0]},
        "clocks": summarize_clocks(samples),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(seq, qual, offs, budget_s=args.cpu_budget)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(seq, qual, offs, budget_s=20.0):
    """The reference's own CPU implementation of the same step on a bounded sample of the same reads:
    FillQualScores + SortByQualScores + PrepareSortedBatch (TBB, all host cores) then ClusterSortedReads -x sahlin
    (single-threaded, as in the reference).  Falls back to the C restatement when oracle/_ref is absent."""
    import oracle
    from isonclust2_b200 import pemp

    n = len(offs) - 1
    if oracle.have_ref():
        R = oracle.ref()
        kind, cores = "reference", os.cpu_count()

        def run(m):
            b = R.sort_and_sketch(seq[:offs[m]], qual[:offs[m]], offs[:m + 1], K, W_, MIN_QUAL)
            b.cluster(0)
    else:
        P = oracle.port()
        kind, cores = "port", 1

        def run(m):
            s, e = P.fill_scores(qual, offs[:m + 1], K)
            o = P.sort_order(s)
            sk = P.sketch_batch(seq, qual, offs[:m + 1], K, W_)
            P.cluster(seq, offs[:m + 1], sk, s, e, o, K, W_, 0, pemp.load_p_shared(K, W_))
    sample = min(n, 1000)
    t0 = time.perf_counter()
    run(sample)
    dt = time.perf_counter() - t0
    if dt < budget_s / 3 and sample < n:  # grow the sample towards the time budget (once)
        sample = int(min(n, sample * min(10.0, budget_s / max(dt, 1e-3) / 1.5)))
        t0 = time.perf_counter()
        run(sample)
        dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "reads/s", "cores": cores, "kind": kind,
            "sample": f"first {sample} reads of the same workload as one batch, {dt:.2f} s "
                      f"(sort+sketch on {cores} threads, cluster single-threaded like the reference)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from isonclust2_b200 import synth

    n_reads = min(args.reads, 20000)
    seq, qual, offs = synth.make_reads(n_reads, seed=100, err_set=(0.08,))
    vals, cb = [], None
    for i in range(min(args.warmup, 1) + args.steps):
        cb = cpu_baseline(seq, qual, offs, budget_s=args.cpu_budget)
        if i >= min(args.warmup, 1):
            vals.append(cb["value"])
    v = float(np.mean(vals)) if vals else cb["value"]
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32/i32 (+f64 quality chains)", "data": "synthetic",
            "config": {"workload": "configs[1] read set, bounded sample on the host cores: sort + cluster -x sahlin"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=100_000, help="reads per GPU")
    ap.add_argument("--batch-max-seq", type=int, default=BATCH_MAX_SEQ)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
