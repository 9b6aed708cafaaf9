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

K, W_, MIN_QUAL = 11, 15, 7.0
BATCH_KB, BATCH_MAX_SEQ = 50000, 30000  # CmdArgs defaults (args.h:14-15)
MODE_IDS = {"sahlin": 0, "fast": 1, "furious": 2}
DIGEST_FILE = os.path.join(ROOT, ".bench_reference_digest.json")  # written by --impl reference, read by the GPU arm


def metric_name(mode, tree=False):
    return f"reads/sec clustered ({mode}" + (", batches merged as a tree)" if tree else ")")


def result_digest(n_clusters, cluster, strand, aln_invoked):
    """sha256 over what a clustering run decides: clusters per batch, (cluster, strand) per sorted slot, ALN_INVOKED."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(n_clusters, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(cluster, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(strand, dtype=np.int32).tobytes())
    h.update(np.int64(aln_invoked).tobytes())
    return h.hexdigest()


def digest_key(reads, seed, mode, batch_max_seq, n_genes):
    return f"reads={reads} seed={seed} genes={n_genes} k={K} w={W_} mode={mode} B={BATCH_KB} M={batch_max_seq} minq={MIN_QUAL}"


# Ceiling of the alignment kernel (integer class): SMs x sub-partitions x 16 ALU lanes per clock (every packed min / max, PRMT and
# add of the kernel issues at 2 clocks per warp instruction and sub-partition: measured, tools/microbench/pipe_rates.cu) x GHz /
# ALU-pipe instructions per cell (212 per 4 steps x 8 rows of a lane in the SASS of icb_k_align16w_dp's inner loop = 6.6) = 2820
ALU_INSTR_PER_CELL = 212 / 32
INT_PIPE_PEAK_GCUPS = 148 * 4 * 16 * 1.965 / ALU_INSTR_PER_CELL
KERNELS = ("sketch_reads", "score_reads", "hpc_error_rate", "index_query", "index_query_pool", "index_query_full", "index_insert", "align_dp",
           "align_trace", "align_dp32", "classify_errors", "index_rehash", "sketch_reads_long", "sketch_reads_cta", "gather_minimizers",
           "dl_init", "dl_decide", "dl_order", "dl_order_large", "dl_commit", "dl_gather", "dl_finalize")


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


def accepted_pairs(res, err, limit):
    """(read, representative, strand, gap_open, window limit) of the reads that JOINED a cluster in `res`, batch by batch — real
    read / representative pairs with the step's length distribution, replayed through icb_align_pairs for the DP roofline."""
    from isonclust2_b200 import api
    order, cl, st = res["order"], res["cluster"], res["strand"]
    out = []
    for (lo, hi) in res["batches"]:
        c = cl[lo:hi]
        live = np.flatnonzero(c >= 0)
        if not len(live):
            continue
        ids, first = np.unique(c[live], return_index=True)
        founder_slot = np.full(int(ids.max()) + 1, -1, np.int64)
        founder_slot[ids] = live[first]
        joined = live[founder_slot[c[live]] != live]
        reads = order[lo + joined]
        reps = order[lo + founder_slot[c[joined]]]
        strands = st[lo:hi][joined]
        for r, p, sd in zip(reads.tolist(), reps.tolist(), strands.tolist()):
            e = err[r] + err[p]
            out.append((r, p, sd, api.set_gap_open(e), api.aln_limit(e, K)))
            if len(out) >= limit:
                return np.array(out, dtype=api.PAIR_DT)
    return np.array(out, dtype=api.PAIR_DT)


def run_ours(args):
    share, cores = cpu_share()
    # torchrun exports OMP_NUM_THREADS=1; the synthetic read generator (OpenMP, outside every timed region) may use this
    # rank's share of the cores
    os.environ["OMP_NUM_THREADS"] = str(share)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # N ranks share the node's cores: size this rank's batch workers and helper pool for its share
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

    mode = args.mode
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

    pipe = dict(k=K, w=W_, mode=mode, min_qual=MIN_QUAL, batch_kb=BATCH_KB, batch_max_seq=args.batch_max_seq, chunk_size=args.chunk,
                n_workers=args.workers, tree=args.tree)

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
    digest = result_digest(res["n_clusters"], res["cluster"], res["strand"], res["stats"]["aln_invoked"])
    if args.quick:  # tuning runs: the device-resident number and the loop's counters only
        stop.set()
        tq = torch.tensor([float(np.mean(times))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        ms = float(tq[0])
        if rank == 0:
            print(json.dumps({"quick": True, "value": world * n_reads / (ms * 1e-3), "ms_per_step": ms, "n_gpus": world,
                              "gpu_launches": int(launches), "result_digest": digest, "cluster_stats": res["stats"],
                              "host_cores_per_rank": share, "clocks": summarize_clocks(samples),
                              "env": {k: v for k, v in os.environ.items() if k.startswith("ICB_")}}), flush=True)
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- kernel-level passes (not part of the timed steps): event pairs around every launch, ONE stream, so the
    # per-kernel durations do not overlap
    nf, nr = rs.sketch_sizes()
    sk = rs.sketch_download(hpc=False)
    hpc_total = int(sk["hpc_len"].sum())
    ctx.timing(True)
    # (1) scores + sketches of the whole read set
    ctx.timing_reset()
    l2_flush.add_(1)
    barrier()
    rs.score_async(K)
    rs.sketch_async(K, W_, skip=api.quality_filtered(res["err"], MIN_QUAL))
    ctx.sync()
    kt = {n: ctx.kernel_time(n) for n in KERNELS}
    # (2) the first batches on a single worker stream: index build / probe kernels, alignment launches as the loop issues them
    ser_batches = res["batches"][:max(1, min(args.serial_batches, len(res["batches"])))]
    ser_bounds = np.array([a for a, _ in ser_batches] + [ser_batches[-1][1]], dtype=np.int64)
    ctx.timing_reset()
    barrier()
    t0 = time.perf_counter()
    _, _, _, ser_stats = api.cluster_batches(rs, res["order"], ser_bounds, K, W_, n_workers=1, mode=mode, min_qual=MIN_QUAL,
                                             chunk_size=args.chunk)
    ctx.sync()
    ser_wall = time.perf_counter() - t0
    ks = {n: ctx.kernel_time(n) for n in KERNELS}
    ser_reads = int(ser_bounds[-1] - ser_bounds[0])
    ser_sel = res["order"][ser_bounds[0]:ser_bounds[-1]]
    ser_mins = int((sk["min_offs"][1:] - sk["min_offs"][:-1])[ser_sel].sum() + (sk["rev_offs"][1:] - sk["rev_offs"][:-1])[ser_sel].sum())
    # (3) alignment kernels saturated: the step's accepted (read, representative) pairs replayed in full launches
    dp_sat = None
    if mode != "fast":
        res_b = res if not args.tree else api.run_pipeline(ctx, None, None, None, rs=rs, **{**pipe, "tree": False})
        pairs = accepted_pairs(res_b, res["err"], args.replay_pairs)
        if len(pairs):
            lens = np.diff(rs.offs)
            cells = int((lens[pairs["read"]].astype(np.int64) * lens[pairs["rep"]].astype(np.int64)).sum())
            rs.align_pairs(pairs[:256], K)  # workspace growth outside the timed launches
            ctx.timing_reset()
            l2_flush.add_(1)
            barrier()
            rs.align_pairs(pairs, K)
            ctx.sync()
            n_dp, ms_dp = ctx.kernel_time("align_dp")
            n_tr, ms_tr = ctx.kernel_time("align_trace")
            dp_sat = {"pairs": int(len(pairs)), "cells": cells, "dp_launches": n_dp, "dp_ms": ms_dp, "trace_ms": ms_tr,
                      "gcups": cells / (ms_dp * 1e-3) / 1e9 if ms_dp > 0 else 0.0}
    ctx.timing(False)
    rs.close()

    # ---------------- e2e through the C ABI with host buffers
    for _ in range(1):
        api.run_pipeline(ctx, seq, qual, offs, **pipe)
    e2e_times = []
    res_e2e = None
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        res_e2e = api.run_pipeline(ctx, seq, qual, offs, **pipe)
        ctx.sync()
        barrier()
        e2e_times.append((time.perf_counter() - t0) * 1e3)
    stop.set()
    th.join(timeout=2)
    digest_e2e = result_digest(res_e2e["n_clusters"], res_e2e["cluster"], res_e2e["strand"], res_e2e["stats"]["aln_invoked"])

    ms = float(np.mean(times))
    ms_e2e = float(np.mean(e2e_times))
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    peak, peak_src = measured_peak()
    # sketch kernel: algorithmic bytes per launch / measured launch time (HBM class)
    alg_bytes = 2 * n_bases + 2 * hpc_total + 12 * (nf + nr) + 24 * n_reads
    n_l, tot_ms = kt["sketch_reads"]
    k_ms = tot_ms / max(n_l, 1)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    t_ratio, t_src = sketch_traffic_ratio()
    roofline_hbm = {"bound": "hbm", "kernel": "icb_k_sketch_warp (sketch_reads)", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": int(t_ratio * alg_bytes) if t_ratio else None,
                    "traffic_source": t_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms}
    # index probe kernel (SURVEY 8d bytes: minimizers in 12 B + bucket heads 8 B per probe, 8 B per posting node walked —
    # twice: counts, then the bitmaps of the relevant candidates —, 16 B per candidate record out)
    q_n, q_ms = ks["index_query_pool"]
    probe_bytes = 20 * ser_mins + 2 * 8 * ser_stats["hits"] + 16 * ser_stats["cands"]
    index_probe = {"bound": "hbm", "kernel": "icb_k_query, pool mode (index_query_pool)", "launches": q_n, "kernel_ms": q_ms,
                   "reads": ser_reads, "hits": ser_stats["hits"], "candidates": ser_stats["cands"],
                   "algorithmic_bytes": probe_bytes, "achieved": probe_bytes / (q_ms * 1e-3) / 1e9 if q_ms > 0 else 0.0, "peak": peak,
                   "unit": "GB/s", "frac": (probe_bytes / (q_ms * 1e-3) / 1e9 / peak) if q_ms > 0 else 0.0,
                   "hits_per_s": ser_stats["hits"] / (q_ms * 1e-3) if q_ms > 0 else 0.0,
                   "note": "latency-bound posting-chain gathers; single worker stream, first batches of the step"}
    st = res["stats"]
    clustered = int((res["cluster"] >= 0).sum())
    eff_gcups = st["cells_aligned"] / (ms * 1e-3) / 1e9
    if mode == "fast":
        roofline = dict(index_probe)
    else:
        # the dominant kernel: icb_k_align16w_dp, integer-ALU class.  achieved = cells of the saturated replay / summed DP launch
        # time on one stream; ceiling = ALU-pipe issue: 15 ALU-pipe instructions per packed update of 2 cells
        roofline = {"bound": "int-alu", "kernel": "icb_k_align16w_dp (align_dp)", "achieved": dp_sat["gcups"] if dp_sat else 0.0,
                    "peak": INT_PIPE_PEAK_GCUPS, "peak_source": "148 SM x 4 x 16 ALU lanes x 1.965 GHz / 6.6 ALU-pipe instr per cell (SASS; issue rates measured, tools/microbench/pipe_rates.cu)",
                    "unit": "GCUPS", "frac": (dp_sat["gcups"] / INT_PIPE_PEAK_GCUPS) if dp_sat else 0.0, "traffic": None,
                    "saturated_replay": dp_sat,
                    "in_step": {"cells_per_step": st["cells_aligned"], "pairs_per_step": st["pairs_aligned"],
                                "effective_gcups": eff_gcups, "effective_frac": eff_gcups / INT_PIPE_PEAK_GCUPS,
                                "aln_invoked": st["aln_invoked"], "aln_calls_ref": st["aln_calls_ref"]}}

    parity, parity_src = None, None
    if rank == 0:
        try:
            ref = json.load(open(DIGEST_FILE))
            key = digest_key(n_reads, 100, mode, args.batch_max_seq, max(1, n_reads // 25)) + (" tree" if args.tree else "")
            if key in ref:
                parity = bool(ref[key]["digest"] == digest and digest_e2e == digest)
                parity_src = f"{os.path.basename(DIGEST_FILE)} written by `bench.py --impl reference` ({ref[key].get('when', '?')})"
        except Exception:
            pass

    line = {
        "metric": metric_name(mode, args.tree), "value": world * n_reads / (ms * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u32/i32 (+f64 quality chains)", "data": "synthetic",
        "config": {**workload_config(n_reads, args.batch_max_seq, mode, args.tree),
                   "reads_per_gpu": n_reads, "bases_per_gpu": n_bases, "batches": len(res["batches"]),
                   "clusters": int(sum(res["n_clusters"])), "reads_clustered": clustered,
                   "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"reads sharded x{world}, no collective",
                   "host_cores": cores, "host_cores_per_rank": share},
        "e2e": {"value": world * n_reads / (ms_e2e * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * n_bases + 8 * (n_reads + 1) + n_reads, "d2h_bytes_per_step": 16 * n_reads + 24 * n_reads,
                "phases_ms_last_step": {k: round(v, 1) for k, v in res_e2e["phases_ms"].items()}},
        "phases_ms_last_step": {k: round(v, 1) for k, v in res["phases_ms"].items()},
        "gpu_launches": int(launches),
        "result_digest": digest, "result_digest_e2e": digest_e2e,
        "parity_vs_reference": parity, "parity_source": parity_src,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "index_probe": index_probe,
        "cluster_stats": st,
        "kernels_serialized": {"what": f"one worker stream: scores + sketches of all reads; cluster of the first {len(ser_batches)} batch(es) "
                                       f"({ser_reads} reads, {ser_wall * 1e3:.0f} ms wall)",
                               "ms": {n: round(v[1], 3) for n, v in {**kt, **{k2: v2 for k2, v2 in ks.items() if v2[0]}}.items() if v[0]},
                               "launches": {n: v[0] for n, v in {**kt, **{k2: v2 for k2, v2 in ks.items() if v2[0]}}.items() if v[0]}},
        "clocks": summarize_clocks(samples),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline_subprocess(args)
        # the same bounded sample through the GPU path: the two must decide identically
        if cb.get("digest"):
            try:
                n_s, m_s = cb["sample_reads"], cb["sample_batch_max_seq"]
                sseq, squal, soffs = synth.make_reads(n_s, seed=100, err_set=(0.08,), n_genes=max(1, n_reads // 25))
                sres = api.run_pipeline(ctx, sseq, squal, soffs, **{**pipe, "batch_max_seq": m_s})
                cb["gpu_digest"] = result_digest(sres["n_clusters"], sres["cluster"], sres["strand"], sres["stats"]["aln_invoked"])
                cb["parity"] = bool(cb["gpu_digest"] == cb["digest"])
            except Exception as ex:  # noqa: BLE001
                cb["parity"] = None
                cb["parity_error"] = repr(ex)[:200]
        line["cpu_baseline"] = cb
        if line["parity_vs_reference"] is None and cb.get("parity") is not None:
            line["parity_vs_reference"] = cb["parity"]
            line["parity_source"] = "cpu_baseline sample (same reads and batch bounds through both implementations)"
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _ref_cluster_worker(args):
    """One `isONclust2 cluster -x <mode>` job: runs in a forked child on the parent's sorted batch (copy-on-write)."""
    lo, hi, nr, mode_id = args
    t0 = time.perf_counter()
    rc = _REF_STATE["batch"].slice(lo, hi, nr).cluster(mode_id)
    return rc["n_clusters"], rc["sizes"], rc["ids"], rc["strands"], rc["n_aligned"], time.perf_counter() - t0


_REF_STATE = {}


def _ref_tree_leaf(args):
    lo, hi, nr, mode_id, path = args
    b = _REF_STATE["batch"].slice(lo, hi, nr)
    rc = b.cluster(mode_id)
    b.save(path)
    return int(rc["n_aligned"])


def _ref_tree_merge(args):
    import oracle
    left, right, out, mode_id = args
    R = oracle.ref()
    a, b = R.load_batch(left), R.load_batch(right)
    rc = a.cluster(mode_id, right=b)
    a.save(out)
    return int(rc["n_aligned"])


def reference_tree_step(seq, qual, offs, batch_max_seq, cores, mode):
    """The reference's pipeline for configs[4] on the host cores: `sort`, one single-threaded `cluster` job per batch, then the
    merge tree level by level (`cluster -l A -r B`), `cores` jobs at a time, every job reading and writing `.cer` files as the
    pipeline does (README.md:105-117).  Returns (seconds, clusters, kind, digest, n_batches)."""
    import multiprocessing as mp
    import shutil
    import tempfile

    import oracle
    from isonclust2_b200 import api

    n = len(offs) - 1
    R = oracle.ref()
    tmp = tempfile.mkdtemp(prefix="icb_ref_tree_")
    try:
        t0 = time.perf_counter()
        batch = R.sort_and_sketch(seq, qual, offs, K, W_, MIN_QUAL)
        _, _, order = batch.scores()
        batches = api.split_batches(order, np.diff(offs), BATCH_KB, batch_max_seq)
        _REF_STATE["batch"] = batch
        mid = MODE_IDS[mode]
        nal = 0
        with mp.get_context("fork").Pool(min(cores, max(1, len(batches)))) as pool:
            paths = [os.path.join(tmp, f"c_{i}.cer") for i in range(len(batches))]
            nal += sum(pool.map(_ref_tree_leaf, [(lo, hi - 1, i, mid, paths[i]) for i, (lo, hi) in enumerate(batches)], chunksize=1))
            level = list(paths)
            while len(level) > 1:  # adjacent nodes pair up, an odd last node is carried up
                jobs, nxt = [], []
                for i in range(0, len(level) - 1, 2):
                    out = level[i] + ".m"
                    jobs.append((level[i], level[i + 1], out, mid))
                    nxt.append(out)
                if len(level) % 2:
                    nxt.append(level[-1])
                nal += sum(pool.map(_ref_tree_merge, jobs, chunksize=1))
                level = nxt
        _REF_STATE.clear()
        root = R.load_batch(level[0])
        dt = time.perf_counter() - t0
        rc = root.cluster_lists()
        slot_of = np.empty(n, np.int64)
        slot_of[order] = np.arange(n)
        cl, st = np.full(n, -1, np.int32), np.zeros(n, np.int32)
        pos = 0
        for ci, sz in enumerate(rc["sizes"]):
            m = rc["ids"][pos + 1:pos + sz]
            cl[slot_of[m]] = ci
            st[slot_of[m]] = rc["strands"][pos + 1:pos + sz]
            pos += sz
        return dt, len(rc["sizes"]), "reference", result_digest([len(rc["sizes"])], cl, st, nal), len(batches)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_step(seq, qual, offs, batch_max_seq, cores, mode):
    """The reference's own CPU implementation of one step over the given reads: `sort` (FillQualScores + SortByQualScores +
    PrepareSortedBatch, TBB on all host cores), batches by the main.cpp:149-199 rule, then one single-threaded
    `cluster -x <mode>` job per batch, `cores` jobs at a time — how the reference's pipeline runs it (README.md:105-117).
    Returns (seconds, clusters, kind, digest, n_batches)."""
    import multiprocessing as mp

    import oracle
    from isonclust2_b200 import api, pemp

    n = len(offs) - 1
    t0 = time.perf_counter()
    if oracle.have_ref():
        R = oracle.ref()
        batch = R.sort_and_sketch(seq, qual, offs, K, W_, MIN_QUAL)
        _, _, order = batch.scores()
        batches = api.split_batches(order, np.diff(offs), BATCH_KB, batch_max_seq)
        _REF_STATE["batch"] = batch
        jobs = [(lo, hi - 1, i, MODE_IDS[mode]) for i, (lo, hi) in enumerate(batches)]
        if len(jobs) == 1:
            out = [_ref_cluster_worker(jobs[0])]
        else:
            with mp.get_context("fork").Pool(min(cores, len(jobs))) as pool:
                out = pool.map(_ref_cluster_worker, jobs, chunksize=1)
        _REF_STATE.clear()
        dt = time.perf_counter() - t0
        # per sorted slot (cluster, strand) from the member lists (element 0 of a cluster is the representative's copy)
        slot_of = np.empty(n, np.int64)
        slot_of[order] = np.arange(n)
        cl, st = np.full(n, -1, np.int32), np.zeros(n, np.int32)
        for (_, sizes, ids, strands, _, _) in out:
            pos = 0
            for ci, sz in enumerate(sizes):
                m = ids[pos + 1:pos + sz]
                cl[slot_of[m]] = ci
                st[slot_of[m]] = strands[pos + 1:pos + sz]
                pos += sz
        ncl = [o[0] for o in out]
        dg = result_digest(ncl, cl, st, sum(int(o[4]) for o in out))
        return dt, int(sum(ncl)), "reference", dg, len(batches)
    # no compiled reference on this box: the C restatement, one thread
    P = oracle.port()
    s, e = P.fill_scores(qual, offs, K)
    o = P.sort_order(s)
    sk = P.sketch_batch(seq, qual, offs, K, W_)
    status = sk["status"].copy()
    status[api.quality_filtered(e, MIN_QUAL).astype(bool)] = 0
    batches = api.split_batches(o, np.diff(offs), BATCH_KB, batch_max_seq)
    cl, st, ncl, nal = np.full(n, -1, np.int32), np.zeros(n, np.int32), [], 0
    for (lo, hi) in batches:
        pc = P.cluster(seq, offs, sk, s, e, o[lo:hi], K, W_, MODE_IDS[mode], pemp.load_p_shared(K, W_), status=status)
        cl[lo:hi], st[lo:hi] = pc["cluster"], pc["strand"]
        ncl.append(pc["n_clusters"])
        nal += pc["n_aligned"]
    return time.perf_counter() - t0, int(sum(ncl)), "port", result_digest(ncl, cl, st, nal), len(batches)


def run_reference(args, quiet=False):
    """--impl reference: the reference's CPU path on the host cores, on the SAME reads and batch bounds as the GPU arm
    (same generator, seed, k / w / mode, split rule).  --sample-reads bounds it to the first reads of that read set (the
    cpu_baseline leg).  Prints the result digest and leaves it in DIGEST_FILE for the GPU arm's parity_vs_reference."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import oracle
    from isonclust2_b200 import synth

    cores = os.cpu_count() or 1
    mode = args.mode
    n_genes = max(1, args.reads // 25)
    n_reads = min(args.reads, args.sample_reads) if args.sample_reads > 0 else args.reads
    if not oracle.have_ref():
        n_reads = min(n_reads, 3000)  # the scalar port is ~20x slower than the compiled reference
    os.environ["OMP_NUM_THREADS"] = str(cores)
    seq, qual, offs = synth.make_reads(n_reads, seed=100, err_set=(0.08,), n_genes=n_genes)
    for _ in range(args.warmup):  # a CPU run has nothing to warm but the page cache: one small batch per warm-up step
        m = min(n_reads, 500)
        reference_step(seq[:offs[m]], qual[:offs[m]], offs[:m + 1], m, 1, mode)
    step_fn = reference_tree_step if (args.tree and oracle.have_ref()) else reference_step
    secs, kind, ncl, dg, nb = [], "port", 0, None, 0
    # one step of the full configuration is minutes of CPU work on every core: steps stop once a minute has been spent
    n_want = min(max(args.steps, 1), 3)
    while len(secs) < n_want:
        dt, ncl, kind, dg, nb = step_fn(seq, qual, offs, args.batch_max_seq, cores, mode)
        secs.append(dt)
        if sum(secs) > 60.0:
            break
    n_steps = len(secs)
    dt = float(np.mean(secs))
    v = n_reads / dt
    used = min(cores, nb) if kind == "reference" else 1
    full = n_reads == args.reads
    cb = {"value": v, "unit": "reads/s", "cores": cores if kind == "reference" else 1, "kind": kind,
          "sample": (f"the whole workload ({n_reads} reads, " if full else f"the first {n_reads} of the workload's {args.reads} reads (") +
                    f"{'' if full else 'same gene set, '}{nb} batches by the main.cpp:149-199 rule, M={args.batch_max_seq}): sort + sketch on "
                    f"{cores} TBB threads, then {used} single-threaded `cluster -x {mode}` jobs in parallel; {dt:.1f} s per step, "
                    f"{ncl} clusters; {n_steps} step(s) timed" + (f" of the {args.steps} requested" if n_steps != args.steps else ""),
          "digest": dg, "sample_reads": n_reads, "sample_batch_max_seq": args.batch_max_seq}
    line = {"impl": "reference", "metric": metric_name(mode, args.tree), "value": v, "unit": "reads/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": n_steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32/i32 (+f64 quality chains)", "data": "synthetic",
            "config": workload_config(args.reads, args.batch_max_seq, mode, args.tree), "result_digest": dg,
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if full and dg:
        try:
            ref = json.load(open(DIGEST_FILE)) if os.path.exists(DIGEST_FILE) else {}
            ref[digest_key(n_reads, 100, mode, args.batch_max_seq, n_genes) + (" tree" if args.tree else "")] = {"digest": dg, "clusters": ncl, "kind": kind,
                                                                                "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
            json.dump(ref, open(DIGEST_FILE, "w"))
        except Exception:
            pass
    if not quiet:
        print(json.dumps(line), flush=True)
    return line


def cpu_baseline_subprocess(args):
    """cpu_baseline of the main line: the reference arm on a bounded sample (the first reads of the same read set, smaller
    batches so that every core has one), in a fresh process (no CUDA state is forked)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0", "--reads", str(args.reads),
           "--mode", args.mode, "--sample-reads", str(min(args.reads, args.cpu_baseline_reads)),
           "--batch-max-seq", str(args.cpu_baseline_batch_reads)] + (["--tree"] if args.tree else [])
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"value": None, "unit": "reads/s", "cores": 0, "kind": "failed", "sample": (r.stderr or "no output")[-300:]}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": "reads/s", "cores": 0, "kind": "failed", "sample": repr(ex)[:300]}


def workload_config(reads, batch_max_seq, mode, tree=False):
    return {"workload": f"configs[1] read set: {reads} synthetic ONT cDNA reads per GPU (~1.3 kb mean, 8% error), k={K} w={W_}, "
                        f"sort + cluster -x {mode} of every batch (B={BATCH_KB} kb, M={batch_max_seq})" +
                        (", then the batches merged as a tree (`cluster -l A -r B` level by level: the shape of configs[4])" if tree else "")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU (configs[1]: 1M)")
    ap.add_argument("--mode", default="sahlin", choices=["sahlin", "fast", "furious"], help="cluster -x (configs[2] = fast, configs[3] = sahlin)")
    ap.add_argument("--workers", type=int, default=0, help="batches in flight per GPU (0 = library default)")
    ap.add_argument("--sample-reads", type=int, default=0, help="--impl reference: only the first reads of the read set (0 = all)")
    ap.add_argument("--cpu-baseline-reads", type=int, default=80000, help="cpu_baseline leg: reads of its sample")
    ap.add_argument("--cpu-baseline-batch-reads", type=int, default=5000, help="cpu_baseline leg: reads per batch of its sample")
    ap.add_argument("--batch-max-seq", type=int, default=BATCH_MAX_SEQ)
    ap.add_argument("--serial-batches", type=int, default=2, help="batches of the single-stream kernel timing pass")
    ap.add_argument("--replay-pairs", type=int, default=60000, help="pairs of the saturated alignment replay")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tree", action="store_true", help="configs[4] shape: the clustered batches are merged as a tree (whole pipeline)")
    ap.add_argument("--quick", action="store_true", help="tuning: device-resident steps only, short line (not the bench contract)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
