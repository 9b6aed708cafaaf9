"""Host-side mirror of the reference's hot-path interface on top of the C ABI.

Names follow the reference (FillQualScores / SortByQualScores / PrepareSortedBatch, qualscore.h:14-24) so the
parity tests read like the reference's own.  Everything numeric happens in libisonclust2_b200.so; this module
only marshals numpy buffers, evaluates the two host-libm expressions the reference evaluates on the host
(InitQualTab's pow, the -10*log10(err) quality filter) and orders reads.
"""
import ctypes as C
import math

import numpy as np

from . import lib as _lib

MIN_DT = np.dtype([("Min", "<u4"), ("Pos", "<u4"), ("Index", "<u4")])
PAIR_DT = np.dtype([("read", "<i4"), ("rep", "<i4"), ("strand", "<i4"), ("gap_open", "<i4"), ("limit", "<i4")])
ALN_DT = np.dtype([("aligned", "<i4"), ("score", "<i4"), ("end_query", "<i4"), ("end_ref", "<i4"), ("comp_len", "<i4")])


def set_gap_open(e):
    """setGapOpen (cluster.cpp:425-440)."""
    return 5 if e <= 0.01 else 4 if e <= 0.04 else 3 if e <= 0.1 else 2


def aln_limit(e, k):
    """floor((1.0 - e) * kmerSize) of getAlnRatio (cluster.cpp:446), as the integer the device compares with."""
    return int(max(-1, min(k + 1, math.floor((1.0 - e) * k))))


class IcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def init_qual_tab():
    """InitQualTab (qualscore.cpp:156-169): host libm pow, capped at 0.79433."""
    t = np.zeros(129)
    _lib.load().icb_init_qual_tab(t.ctypes.data, 0)
    return t


def init_qual_tab_nomin():
    """InitQualTabNomin (qualscore.cpp:171-180)."""
    t = np.zeros(129)
    _lib.load().icb_init_qual_tab(t.ctypes.data, 1)
    return t


def quality_filtered(err, min_qual):
    """(-10*log10(err)) <= minQual with the host libm, like qualscore.cpp:56 / cluster.cpp:157."""
    err = np.ascontiguousarray(err, dtype=np.float64)
    out = np.zeros(len(err), dtype=np.uint8)
    rc = _lib.load().icb_quality_filter(err.ctypes.data, len(err), float(min_qual), out.ctypes.data)
    if rc:
        raise IcbError(rc, "icb_quality_filter")
    return out


def sort_by_qual_scores(score):
    """SortByQualScores (qualscore.cpp:138-145): std::stable_sort, descending."""
    score = np.ascontiguousarray(score, dtype=np.float64)
    order = np.empty(len(score), dtype=np.int32)
    rc = _lib.load().icb_sort_by_score(score.ctypes.data, len(score), order.ctypes.data)
    if rc:
        raise IcbError(rc, "icb_sort_by_score")
    return order


class ClusterParams(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("window_size", C.c_int32), ("mode", C.c_int32), ("min_shared", C.c_int32),
                ("mapped_threshold", C.c_double), ("aligned_threshold", C.c_double), ("min_fraction", C.c_double),
                ("min_prob_no_hits", C.c_double), ("min_qual", C.c_double), ("p_shared", C.c_double * 225),
                ("chunk_size", C.c_int32)]


class ClusterStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("chunks", "rounds", "reads_queried", "aln_invoked", "pairs_aligned", "cells_aligned",
                                         "tie_fallbacks", "decisions")] + \
               [(n, C.c_double) for n in ("ms_query", "ms_full", "ms_align", "ms_insert", "ms_host", "ms_cpu", "ms_order")] + \
               [("order_calls", C.c_int64), ("aln_calls_ref", C.c_int64), ("ms_batch_max", C.c_double), ("ms_wall", C.c_double), ("hyp_decisions", C.c_int64),
                ("hits", C.c_int64), ("cands", C.c_int64), ("dl_chunks", C.c_int64), ("dl_fallback_chunks", C.c_int64),
                ("tree_merges", C.c_int64), ("dl_fallback_reasons", C.c_int64),
                ("dl_stop_full", C.c_int64), ("dl_stop_pair_first", C.c_int64), ("dl_stop_pair_again", C.c_int64)]


MODES = {"sahlin": 0, "fast": 1, "furious": 2}


def make_params(k, w, mode="sahlin", min_shared=5, mapped_threshold=0.65, aligned_threshold=0.2, min_fraction=0.8,
                min_prob_no_hits=0.1, min_qual=7.0, chunk_size=0, p_shared=None):
    """CmdArgs defaults of args.h:13-28 + the InitMinSharedMap(k, w) grid."""
    from . import pemp
    P = ClusterParams(kmer_size=k, window_size=w, mode=MODES.get(mode, mode), min_shared=min_shared,
                      mapped_threshold=mapped_threshold, aligned_threshold=aligned_threshold, min_fraction=min_fraction,
                      min_prob_no_hits=min_prob_no_hits, min_qual=min_qual, chunk_size=chunk_size)
    ps = np.ascontiguousarray(pemp.load_p_shared(k, w) if p_shared is None else p_shared, dtype=np.float64).reshape(225)
    for i in range(225):
        P.p_shared[i] = ps[i]
    return P


def cluster_batches(rs, order, bounds, k, w, n_workers=0, **params):
    """One `isONclust2 cluster -x <mode>` per batch order[bounds[b]:bounds[b+1]], several batches in flight on the device.

    Returns (cluster per slot (batch-local ids), strand per slot, clusters per batch, stats summed)."""
    c = rs.ctx
    P = make_params(k, w, **params)
    order = np.ascontiguousarray(order, dtype=np.int32)
    bounds = np.ascontiguousarray(bounds, dtype=np.int64)
    nb = len(bounds) - 1
    cl, st = np.full(len(order), -1, np.int32), np.zeros(len(order), np.int32)
    ncl = np.zeros(max(nb, 1), np.int64)
    stats = ClusterStats()
    c.check(c.L.icb_cluster_batches(rs.h, C.byref(P), order.ctypes.data, bounds.ctypes.data, nb, n_workers, cl.ctypes.data,
                                    st.ctypes.data, ncl.ctypes.data, C.byref(stats)))
    return cl, st, ncl[:nb], {n: getattr(stats, n) for n, _ in ClusterStats._fields_}


def cluster_tree(rs, order, bounds, k, w, n_workers=0, min_cls_size=3, **params):
    """configs[4] in process: every batch clustered, then merged as a tree (`cluster -l A -r B` level by level) on clusterers
    that stay resident on the device.  Returns (cluster per slot in the ROOT, strand per slot, clusters of the root, stats)."""
    c = rs.ctx
    P = make_params(k, w, **params)
    order = np.ascontiguousarray(order, dtype=np.int32)
    bounds = np.ascontiguousarray(bounds, dtype=np.int64)
    nb = len(bounds) - 1
    cl, st = np.full(len(order), -1, np.int32), np.zeros(len(order), np.int32)
    ncl = C.c_int64()
    stats = ClusterStats()
    c.check(c.L.icb_cluster_tree(rs.h, C.byref(P), order.ctypes.data, bounds.ctypes.data, nb, n_workers, min_cls_size, cl.ctypes.data,
                                 st.ctypes.data, C.byref(ncl), C.byref(stats)))
    return cl, st, ncl.value, {n: getattr(stats, n) for n, _ in ClusterStats._fields_}


class Clusterer:
    """ClusterSortedReads (cluster.cpp:67-322) for an initial clustering; CmdArgs defaults of args.h:13-28."""

    def __init__(self, rs, k, w, mode="sahlin", min_shared=5, mapped_threshold=0.65, aligned_threshold=0.2, min_fraction=0.8,
                 min_prob_no_hits=0.1, min_qual=7.0, chunk_size=0, p_shared=None):
        self.rs = rs
        c = rs.ctx
        P = make_params(k, w, mode=mode, min_shared=min_shared, mapped_threshold=mapped_threshold,
                        aligned_threshold=aligned_threshold, min_fraction=min_fraction, min_prob_no_hits=min_prob_no_hits,
                        min_qual=min_qual, chunk_size=chunk_size, p_shared=p_shared)
        h = C.c_void_p()
        c.check(c.L.icb_clusterer_create(rs.h, C.byref(P), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.rs.ctx.L.icb_clusterer_destroy(self.h)
            self.h = None

    def seed(self, rep_reads):
        """Existing clusters (the left batch of a merge): rep_reads[c] represents the next cluster id."""
        c = self.rs.ctx
        rep = np.ascontiguousarray(rep_reads, dtype=np.int32)
        c.check(c.L.icb_clusterer_seed(self.h, rep.ctypes.data, len(rep)))

    def update_representative(self, cluster, rep_read):
        """UpdateMinDB (minimizer.cpp:124-160): read `rep_read` represents `cluster` from now on."""
        c = self.rs.ctx
        c.check(c.L.icb_clusterer_update_representative(self.h, int(cluster), int(rep_read)))

    def cluster(self, order):
        c = self.rs.ctx
        order = np.ascontiguousarray(order, dtype=np.int32)
        cl, st = np.empty(len(order), np.int32), np.empty(len(order), np.int32)
        c.check(c.L.icb_cluster_reads(self.h, order.ctypes.data, len(order), cl.ctypes.data, st.ctypes.data))
        return cl, st

    @property
    def n_clusters(self):
        return self.rs.ctx.L.icb_clusterer_n_clusters(self.h)

    def founders(self):
        f = np.empty(self.n_clusters, np.int32)
        self.rs.ctx.check(self.rs.ctx.L.icb_clusterer_founders(self.h, f.ctypes.data))
        return f

    def stats(self):
        s = ClusterStats()
        self.rs.ctx.check(self.rs.ctx.L.icb_clusterer_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in ClusterStats._fields_}


class Context:
    def __init__(self, device=0):
        self.L = _lib.load()
        h = C.c_void_p()
        rc = self.L.icb_ctx_create(device, C.byref(h))
        if rc != 0:
            raise IcbError(rc, self.L.icb_strerror(rc).decode() + " (icb_ctx_create; is a CUDA device visible?)")
        self.h = h
        self.device = device
        self.qualtab = init_qual_tab()
        self.qualtab_nomin = init_qual_tab_nomin()

    def check(self, rc):
        if rc != 0:
            raise IcbError(rc, self.L.icb_last_error(self.h).decode() or self.L.icb_strerror(rc).decode())

    def close(self):
        if self.h:
            self.L.icb_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self.check(self.L.icb_ctx_sync(self.h))

    @property
    def stream(self):
        return self.L.icb_ctx_stream(self.h)

    def timing(self, enable=True):
        self.check(self.L.icb_ctx_timing(self.h, 1 if enable else 0))

    def timing_reset(self):
        self.check(self.L.icb_ctx_timing_reset(self.h))

    def kernel_time(self, name):
        n, ms = C.c_int64(), C.c_double()
        self.check(self.L.icb_ctx_kernel_time(self.h, name.encode(), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def launch_count(self):
        return self.L.icb_ctx_launch_count(self.h)

    def upload(self, seq, qual, offs, reserve=None):
        return ReadSet(self, seq, qual, offs, reserve=reserve)


class PoaGraph:
    """spoa::Graph + the reference's alignment engine on the HOST (csrc/poa.cpp, bit-exact to the vendored spoa 4.0.9): the
    consensus side of `sort -c <positive>` clustering (consensus.cpp:15-137).  algo: 0 local, 1 global, 2 semi-global."""

    def __init__(self, blob=None):
        self.L = _lib.load()
        self.h = C.c_void_p()
        rc = self.L.icb_poa_graph_create(C.byref(self.h)) if blob is None else self.L.icb_poa_graph_load(blob, len(blob), C.byref(self.h))
        if rc:
            raise IcbError(rc, "cannot create / load a POA graph")

    def close(self):
        if self.h:
            self.L.icb_poa_graph_destroy(self.h)
            self.h = None

    def _check(self, rc, what):
        if rc:
            raise IcbError(rc, what)

    def add(self, seq, weight=1, algo=2):  # AddSeqToGraph
        self._check(self.L.icb_poa_graph_add(self.h, bytes(seq), len(seq), int(weight), int(algo)), "icb_poa_graph_add")

    def consensus(self):  # Graph::GenerateConsensus
        n = C.c_int64()
        self._check(self.L.icb_poa_graph_consensus(self.h, None, 0, C.byref(n)), "icb_poa_graph_consensus")
        buf = C.create_string_buffer(max(1, n.value))
        self._check(self.L.icb_poa_graph_consensus(self.h, buf, n.value, C.byref(n)), "icb_poa_graph_consensus")
        return buf.raw[:n.value]

    def save(self):  # the ConsGs blob of a `.cer`
        n = C.c_int64()
        self._check(self.L.icb_poa_graph_save(self.h, None, 0, C.byref(n)), "icb_poa_graph_save")
        buf = C.create_string_buffer(max(1, n.value))
        self._check(self.L.icb_poa_graph_save(self.h, buf, n.value, C.byref(n)), "icb_poa_graph_save")
        return buf.raw[:n.value]

    @property
    def n_sequences(self):
        n = C.c_int64()
        self._check(self.L.icb_poa_graph_sizes(self.h, C.byref(n), None, None), "icb_poa_graph_sizes")
        return n.value

    def clear(self):
        self._check(self.L.icb_poa_graph_clear(self.h), "icb_poa_graph_clear")


class ReadSet:
    """Device-resident flat reads + their scores and sketches.  reserve = (extra_reads, extra_bases, extra_max_len): room for
    reads appended later with append_from (consensus representatives)."""

    def __init__(self, ctx, seq, qual, offs, reserve=None):
        self.ctx = ctx
        L = ctx.L
        seq = np.ascontiguousarray(seq, dtype=np.uint8) if not isinstance(seq, np.ndarray) else seq
        qual = np.ascontiguousarray(qual, dtype=np.uint8) if not isinstance(qual, np.ndarray) else qual
        self.offs = np.ascontiguousarray(offs, dtype=np.int64)
        self.n = len(self.offs) - 1
        self.n_bases = int(self.offs[-1])
        h = C.c_void_p()
        if reserve is None:
            ctx.check(L.icb_readset_create(ctx.h, seq.ctypes.data, qual.ctypes.data, self.offs.ctypes.data, self.n, C.byref(h)))
        else:
            er, eb, el = (int(v) for v in reserve)
            ctx.check(L.icb_readset_create_reserve(ctx.h, seq.ctypes.data, qual.ctypes.data, self.offs.ctypes.data, self.n, er, eb, el, C.byref(h)))
        self.h = h

    def append_from(self, src, read, raw_err, raw_score, hpc_err):
        """Read `read` of the sketched read set `src` becomes the next read of this one, with the given error rates
        (UpdateClusterConsensus' weighted means, consensus.cpp:57-63).  Returns its index."""
        idx = C.c_int64()
        self.ctx.check(self.ctx.L.icb_readset_append_from(self.h, src.h, int(read), float(raw_err), float(raw_score), float(hpc_err), C.byref(idx)))
        return idx.value

    def set_errors(self, raw_err=None, raw_score=None, hpc_err=None):
        """Stored error rates / scores of a loaded batch instead of the computed ones (before the first Clusterer)."""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (raw_err, raw_score, hpc_err)]
        self.ctx.check(self.ctx.L.icb_readset_set_errors(self.h, *[None if a is None else a.ctypes.data for a in arrs]))

    def close(self):
        if self.h:
            self.ctx.L.icb_readset_destroy(self.h)
            self.h = None

    # FillQualScores
    def score_async(self, k):
        c = self.ctx
        c.check(c.L.icb_score_reads(self.h, k, c.qualtab.ctypes.data, c.qualtab_nomin.ctypes.data))

    def scores(self, score=None, err=None):
        c = self.ctx
        score = np.empty(self.n) if score is None else score
        err = np.empty(self.n) if err is None else err
        c.check(c.L.icb_scores_download(self.h, score.ctypes.data, err.ctypes.data))
        return score, err

    def score(self, k):
        self.score_async(k)
        return self.scores()

    # PrepareSortedBatch body
    def sketch_async(self, k, w, skip=None):
        c = self.ctx
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
        c.check(c.L.icb_sketch_reads(self.h, k, w, c.qualtab_nomin.ctypes.data, skip.ctypes.data if skip is not None else None))

    def sketch_sizes(self):
        a, b = C.c_int64(), C.c_int64()
        self.ctx.check(self.ctx.L.icb_sketch_sizes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sketch_download(self, hpc=True):
        c = self.ctx
        nf, nr = self.sketch_sizes()
        n = self.n
        status, hpc_len = np.empty(n, np.int32), np.empty(n, np.int32)
        hpc_err = np.empty(n)
        hs = np.empty(self.n_bases, np.uint8) if hpc else None
        hq = np.empty(self.n_bases, np.uint8) if hpc else None
        mins, rmins = np.empty(nf, MIN_DT), np.empty(nr, MIN_DT)
        mo, ro = np.empty(n + 1, np.int64), np.empty(n + 1, np.int64)
        c.check(c.L.icb_sketch_download(self.h, status.ctypes.data, hpc_len.ctypes.data, hpc_err.ctypes.data,
                                        hs.ctypes.data if hpc else None, hq.ctypes.data if hpc else None,
                                        mins.ctypes.data, mo.ctypes.data, rmins.ctypes.data, ro.ctypes.data))
        return dict(status=status, hpc_len=hpc_len, hpc_err=hpc_err, hpc_seq=hs, hpc_qual=hq, mins=mins, min_offs=mo,
                    rev_mins=rmins, rev_offs=ro)

    def align_pairs(self, pairs, k):
        """pairs: PAIR_DT array (indices into this read set).  Returns ALN_DT array."""
        c = self.ctx
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DT)
        out = np.zeros(len(pairs), ALN_DT)
        c.check(c.L.icb_align_pairs(self.h, pairs.ctypes.data, len(pairs), k, out.ctypes.data))
        return out

    def sketch(self, k, w, skip=None, hpc=True):
        self.sketch_async(k, w, skip)
        return self.sketch_download(hpc=hpc)


def split_batches(order, lengths, batch_kb=50000, batch_max_seq=30000):
    """The batch windows of mainSort (main.cpp:149-199): a batch closes with the read that takes it above
    batch_kb*1000 bases or to batch_max_seq reads.  Returns [(start, end_exclusive)] over the sorted order."""
    out, start, bases, cnt = [], 0, 0, 0
    csum = np.cumsum(np.asarray(lengths)[np.asarray(order)])
    n = len(order)
    i = 0
    while i < n:
        # first index where either limit is hit
        base0 = csum[start - 1] if start > 0 else 0
        lim_b = int(np.searchsorted(csum, base0 + batch_kb * 1000, side="right")) if batch_kb > 0 else n
        lim_n = start + batch_max_seq - 1 if batch_max_seq > 0 else n
        end = min(lim_b, lim_n, n - 1)
        if batch_kb <= 0:
            end = n - 1
        out.append((start, end + 1))
        start = end + 1
        i = start
    return out


def run_pipeline(ctx, seq, qual, offs, k=11, w=15, mode="sahlin", min_qual=7.0, batch_kb=50000, batch_max_seq=30000,
                 chunk_size=0, rs=None, keep=False, n_workers=0, tree=False, min_cls_size=3, **cluster_kw):
    """`isONclust2 sort` + one `isONclust2 cluster -x <mode>` per batch, on one device, without the files in between.
    tree=True: the batches are then merged as a tree (`cluster -l A -r B` level by level) on resident clusterers.

    Returns dict(order, batches, cluster (per sorted slot; batch-local ids, or the root's with tree=True), strand,
    n_clusters per batch (tree=True: [clusters of the root]), stats)."""
    own = rs is None
    if own:
        rs = ctx.upload(seq, qual, offs)
    import time
    t = [time.perf_counter()]
    lap = lambda: (t.append(time.perf_counter()), (t[-1] - t[-2]) * 1e3)[1]
    if own:
        ctx.sync()
    phases = {"upload": lap()}
    try:
        rs.score_async(k)
        score, err = rs.scores()
        phases["score"] = lap()
        skip = quality_filtered(err, min_qual)
        order = sort_by_qual_scores(score)
        phases["sort"] = lap()
        rs.sketch_async(k, w, skip=skip)
        lengths = np.diff(rs.offs)
        batches = split_batches(order, lengths, batch_kb, batch_max_seq)
        bounds = np.array([a for a, _ in batches] + [batches[-1][1] if batches else 0], dtype=np.int64)
        phases["sketch_split"] = lap()
        if tree:
            cluster, strand, nroot, tot = cluster_tree(rs, order, bounds, k, w, n_workers=n_workers, min_cls_size=min_cls_size, mode=mode,
                                                       min_qual=min_qual, chunk_size=chunk_size, **cluster_kw)
            ncl = [nroot]
        else:
            cluster, strand, ncl, tot = cluster_batches(rs, order, bounds, k, w, n_workers=n_workers, mode=mode, min_qual=min_qual,
                                                        chunk_size=chunk_size, **cluster_kw)
        phases["cluster"] = lap()
        ncl = [int(x) for x in ncl]
        return dict(order=order, batches=batches, cluster=cluster, strand=strand, n_clusters=ncl, stats=tot, score=score, err=err,
                    phases_ms=phases)
    finally:
        if own and not keep:
            rs.close()
