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
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before torch initialises CUDA: one hardware queue per worker stream

K, W_, MIN_QUAL, MODE = 11, 15, 7.0, "sahlin"
BATCH_KB, BATCH_MAX_SEQ = 50000, 30000  # CmdArgs defaults (args.h:14-15)
METRIC = "reads/sec clustered (sahlin)"
INT_PIPE_PEAK_GCUPS = 148 * 4 * 16 * 1.965 / 7.5  # SMs x sub-partitions x ALU lanes/clk x GHz / ALU instructions per cell = 2482
KERNELS = ("sketch_reads", "score_reads", "hpc_error_rate", "index_query", "index_query_full", "index_insert", "align_dp",
           "align_trace", "align_dp32", "classify_errors", "index_rehash", "sketch_reads_long", "sketch_reads_cta", "gather_minimizers")


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


def sketch_traffic_ratio():
    """DRAM bytes per algorithmic byte of the sketch kernel, from the committed `ncu --set full` capture (newest
    profiles/*_sketch_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch over that launch's algorithmic
    bytes).  bench.py never runs under a profiler itself; the ratio is applied to this run's algorithmic bytes."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_sketch_traffic.json")))
    if not files:
        return None, None
    try:
        return float(json.load(open(files[-1]))["dram_bytes_per_algorithmic_byte"]), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def cpu_share():
    """Host cores this rank may count on: the process's affinity mask split evenly over the ranks of this node."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    return max(2, cores // max(local_world, 1)), cores


def run_ours(args):
    share, cores = cpu_share()
    # torchrun exports OMP_NUM_THREADS=1; the synthetic read generator (OpenMP, outside every timed region) may use this
    # rank's share of the cores
    os.environ["OMP_NUM_THREADS"] = str(share)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # N ranks share the node's cores: size this rank's batch workers and helper pool for its share (library defaults
        # assume the whole machine: 3 workers per core, one helper per core).  Measured on 8 cores, 1 M reads, N=1
        # (taskset): 24 workers 279 k reads/s, the default 48 (35 batches in flight) 233 k reads/s.
        os.environ.setdefault("ICB_HOST_CORES", str(share))
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

    pipe = dict(k=K, w=W_, mode=MODE, min_qual=MIN_QUAL, batch_kb=BATCH_KB, batch_max_seq=args.batch_max_seq, chunk_size=args.chunk,
                n_workers=args.workers)

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
    # one more pass with an event pair around every kernel launch (the library's per-kernel timers): kernel-level
    # durations for the roofline; not part of the timed steps above
    ctx.timing(True)
    ctx.timing_reset()
    l2_flush.add_(1)
    barrier()
    res = api.run_pipeline(ctx, None, None, None, rs=rs, **pipe)
    ctx.sync()
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
        res_e2e = api.run_pipeline(ctx, seq, qual, offs, **pipe)
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
    steps = 1  # the kernel timers cover the single profiling pass
    alg_bytes = 2 * n_bases + 2 * hpc_total + 12 * (nf + nr) + 24 * n_reads
    n_l, tot_ms = kt["sketch_reads"]
    k_ms = tot_ms / max(n_l, 1)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    t_ratio, t_src = sketch_traffic_ratio()
    st = res["stats"]
    dp_ms = kt["align_dp"][1] / steps
    gcups = st["cells_aligned"] / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else 0.0
    clustered = int((res["cluster"] >= 0).sum())

    line = {
        "metric": METRIC, "value": world * n_reads / (ms * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u32/i32 (+f64 quality chains)", "data": "synthetic",
        "config": {**workload_config(n_reads, args.batch_max_seq),
                   "reads_per_gpu": n_reads, "bases_per_gpu": n_bases, "batches": len(res["batches"]),
                   "clusters": int(sum(res["n_clusters"])), "reads_clustered": clustered,
                   "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"reads sharded x{world}, no collective", "host_cores": cores, "host_cores_per_rank": share},
        "e2e": {"value": world * n_reads / (ms_e2e * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * n_bases + 8 * (n_reads + 1) + n_reads, "d2h_bytes_per_step": 16 * n_reads + 24 * n_reads,
                "phases_ms_last_step": {k: round(v, 1) for k, v in res_e2e["phases_ms"].items()}},
        "phases_ms_last_step": {k: round(v, 1) for k, v in res["phases_ms"].items()},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "icb_k_sketch (sketch_reads)", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": int(t_ratio * alg_bytes) if t_ratio else None,
                     "traffic_source": t_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms},
        # integer-ALU class (no "hbm"/"tensor" bound fits): cells per second of the whole step against the ALU-pipe ceiling of
        # the DP update (15 ALU-pipe instructions per 2 cells, 16 lanes/clk per SM sub-partition).
        # `gcups` divides by the SUMMED durations of launches that overlap on 35 streams; `effective_gcups` by the step time.
        "alignment": {"gcups": gcups, "effective_gcups": st["cells_aligned"] / (ms * 1e-3) / 1e9,
                      "int_pipe_peak_gcups": INT_PIPE_PEAK_GCUPS,
                      "effective_frac": st["cells_aligned"] / (ms * 1e-3) / 1e9 / INT_PIPE_PEAK_GCUPS,
                      "cells_per_step": st["cells_aligned"], "pairs_per_step": st["pairs_aligned"],
                      "aln_invoked": st["aln_invoked"], "dp_ms_per_step": dp_ms},
        "cluster_stats": st,
        "kernels_ms_per_step": {n: v[1] / steps for n, v in kt.items() if v[0]},
        "kernel_launches_per_step": {n: v[0] / steps for n, v in kt.items() if v[0]},
        "clocks": summarize_clocks(samples),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _ref_cluster_worker(args):
    """One `isONclust2 cluster -x sahlin` job: runs in a forked child on the parent's sorted batch (copy-on-write)."""
    lo, hi, nr = args
    t0 = time.perf_counter()
    rc = _REF_STATE["batch"].slice(lo, hi, nr).cluster(0)
    return rc["n_clusters"], time.perf_counter() - t0


_REF_STATE = {}


def reference_step(seq, qual, offs, batch_reads, cores):
    """The reference's own CPU implementation of one step over the given reads: `sort` (FillQualScores + SortByQualScores +
    PrepareSortedBatch, TBB on all host cores) then one single-threaded `cluster -x sahlin` job per batch, `cores` jobs at a
    time — how the reference's pipeline runs it (README.md:105-117).  Returns (seconds, clusters, kind)."""
    import multiprocessing as mp

    import oracle
    from isonclust2_b200 import pemp

    n = len(offs) - 1
    bounds = list(range(0, n, batch_reads)) + [n]
    t0 = time.perf_counter()
    if oracle.have_ref():
        R = oracle.ref()
        batch = R.sort_and_sketch(seq, qual, offs, K, W_, MIN_QUAL)
        _REF_STATE["batch"] = batch
        jobs = [(bounds[i], bounds[i + 1] - 1, i) for i in range(len(bounds) - 1)]
        if len(jobs) == 1:
            out = [_ref_cluster_worker(jobs[0])]
        else:
            with mp.get_context("fork").Pool(min(cores, len(jobs))) as pool:
                out = pool.map(_ref_cluster_worker, jobs, chunksize=1)
        _REF_STATE.clear()
        return time.perf_counter() - t0, int(sum(o[0] for o in out)), "reference"
    # no compiled reference on this box: the C restatement, one thread
    P = oracle.port()
    s, e = P.fill_scores(qual, offs, K)
    o = P.sort_order(s)
    sk = P.sketch_batch(seq, qual, offs, K, W_)
    ncl = 0
    for i in range(len(bounds) - 1):
        ncl += P.cluster(seq, offs, sk, s, e, o[bounds[i]:bounds[i + 1]], K, W_, 0, pemp.load_p_shared(K, W_))["n_clusters"]
    return time.perf_counter() - t0, int(ncl), "port"


def run_reference(args, quiet=False):
    """--impl reference: the reference's CPU path on the host cores, on a bounded sample of the same workload
    (same generator and seed, same k / w / mode / batch size; `cores` batches so every core has one cluster job)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import oracle
    from isonclust2_b200 import synth

    cores = os.cpu_count() or 1
    batch_reads = min(args.batch_max_seq, args.ref_batch_reads)
    n_batches = max(1, min(cores, args.reads // batch_reads)) if oracle.have_ref() else 1
    n_reads = n_batches * batch_reads
    seq, qual, offs = synth.make_reads(args.reads, seed=100, err_set=(0.08,))
    seq, qual, offs = seq[:offs[n_reads]], qual[:offs[n_reads]], offs[:n_reads + 1]
    for _ in range(args.warmup):  # a CPU run has nothing to warm but the page cache: one small batch per warm-up step
        m = min(n_reads, 500)
        reference_step(seq[:offs[m]], qual[:offs[m]], offs[:m + 1], m, 1)
    secs, kind, ncl = [], "port", 0
    # one step is about a minute of CPU work on every core: at most 3 are timed, so that any --steps K ends within minutes
    n_steps = min(max(args.steps, 1), 3)
    for _ in range(n_steps):
        dt, ncl, kind = reference_step(seq, qual, offs, batch_reads, cores)
        secs.append(dt)
    dt = float(np.mean(secs))
    v = n_reads / dt
    used = min(cores, n_batches) if kind == "reference" else 1
    cb = {"value": v, "unit": "reads/s", "cores": cores if kind == "reference" else 1, "kind": kind,
          "sample": f"{n_reads} reads of the same workload = {n_batches} batches of {batch_reads}: sort + sketch on {cores} TBB threads, "
                    f"then {used} single-threaded `cluster -x sahlin` jobs in parallel; {dt:.1f} s per step, {ncl} clusters; "
                    f"{n_steps} step(s) timed" + (f" of the {args.steps} requested" if n_steps != args.steps else "")}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": n_steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32/i32 (+f64 quality chains)", "data": "synthetic",
            "config": workload_config(args.reads, args.batch_max_seq),
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not quiet:
        print(json.dumps(line), flush=True)
    return line


def cpu_baseline_subprocess(args):
    """cpu_baseline of the main line: the reference arm on a smaller bounded sample, in a fresh process (no CUDA state
    is forked)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0", "--reads", str(args.reads),
           "--batch-max-seq", str(args.batch_max_seq), "--ref-batch-reads", str(args.cpu_baseline_batch_reads)]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"value": None, "unit": "reads/s", "cores": 0, "kind": "failed", "sample": (r.stderr or "no output")[-300:]}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": "reads/s", "cores": 0, "kind": "failed", "sample": repr(ex)[:300]}


def workload_config(reads, batch_max_seq):
    return {"workload": f"configs[1] read set: {reads} synthetic ONT cDNA reads per GPU (~1.3 kb mean, 8% error), k={K} w={W_}, "
                        f"sort + cluster -x {MODE} of every batch (B={BATCH_KB} kb, M={batch_max_seq})"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU (configs[1]: 1M)")
    ap.add_argument("--workers", type=int, default=0, help="batches in flight per GPU (0 = library default)")
    ap.add_argument("--ref-batch-reads", type=int, default=BATCH_MAX_SEQ, help="--impl reference: reads per batch of its sample")
    ap.add_argument("--cpu-baseline-batch-reads", type=int, default=5000, help="cpu_baseline leg: reads per batch of its sample")
    ap.add_argument("--batch-max-seq", type=int, default=BATCH_MAX_SEQ)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
