"""TEST INFRASTRUCTURE: ctypes front ends of the two CPU checkers.

    oracle.port  — libisonclust2_oracle.so, the plain-C/C++ restatement (always available after build())
    oracle.ref   — _ref/libisonclust2_ref.so, the UNMODIFIED reference compiled from /root/reference
                   (built in the dev container, ships to the GPU box as a prebuilt file)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libisonclust2_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libisonclust2_ref.so")
REF_BIN = os.path.join(_HERE, "_ref", "isONclust2")

MIN_DT = np.dtype([("Min", "<u4"), ("Pos", "<u4"), ("Index", "<u4")])

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_dbl = C.c_double


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _bytes(a):
    return np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a, dtype=np.uint8)


class ClusterParams(C.Structure):
    _fields_ = [("k", _i32), ("w", _i32), ("mode", _i32), ("min_shared", _i32), ("mapped_th", _dbl),
                ("aligned_th", _dbl), ("min_frac", _dbl), ("min_prob_no_hits", _dbl), ("min_qual", _dbl),
                ("p_shared", _dbl * 225)]


class Reads(C.Structure):
    _fields_ = [("n_reads", _i64), ("status", _p), ("score", _p), ("raw_err", _p), ("hpc_err", _p), ("raw_seq", _p),
                ("raw_offs", _p), ("hpc_len", _p), ("mins", _p), ("min_offs", _p), ("rev_mins", _p), ("rev_offs", _p)]


def _mindb_after_updates(fn, cls_mins, cls_offs, upd_cls, upd_mins, upd_offs):
    """AddMinimizers for the clusters, then UpdateMinDB per update; returns (keys sorted, offs, cls)."""
    cm = np.ascontiguousarray(cls_mins, MIN_DT); co = np.ascontiguousarray(cls_offs, np.int64)
    uc = np.ascontiguousarray(upd_cls, np.int32); um = np.ascontiguousarray(upd_mins, MIN_DT)
    uo = np.ascontiguousarray(upd_offs, np.int64)
    cap = len(cm) + len(um) + 1
    keys, offs, cls = np.zeros(cap, np.uint32), np.zeros(cap + 1, np.int64), np.zeros(cap, np.uint32)
    n = fn(_ptr(cm), _ptr(co), len(co) - 1, _ptr(uc), _ptr(um), _ptr(uo), len(uc), _ptr(keys), _ptr(offs), _ptr(cls), cap, cap)
    assert n >= 0, n
    return keys[:n].copy(), offs[:n + 1].copy(), cls[:offs[n]].copy()


class Port:
    """The C restatement."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            raise RuntimeError(f"{PORT_SO} missing: run __graft_entry__.build()")
        L = self.L = C.CDLL(PORT_SO)
        L.orc_calc_qual_score.restype = _dbl
        L.orc_calc_qual_score.argtypes = [_p, _i64, _i32, _p]
        L.orc_calc_error_rate.restype = _dbl
        L.orc_calc_error_rate.argtypes = [_p, _i64, _p]
        L.orc_fill_qual_score.argtypes = [_p, _i64, _i32, _p, _p, _p, _p]
        L.orc_sort_by_score.argtypes = [_p, _i64, _p]
        L.orc_hpc.restype = _i64
        L.orc_hpc.argtypes = [_p, _p, _i64, _p, _p]
        L.orc_revcomp.restype = _i64
        L.orc_revcomp.argtypes = [_p, _i64, _p]
        L.orc_kmer_encode.restype = _i64
        L.orc_kmer_encode.argtypes = [_p, _i64, _i32, _p]
        L.orc_minimizers.restype = _i64
        L.orc_minimizers.argtypes = [_p, _i64, _i32, _i32, _p]
        L.orc_sketch_read.restype = _i32
        L.orc_sketch_read.argtypes = [_p, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]
        L.orc_quality_filtered.restype = _i32
        L.orc_quality_filtered.argtypes = [_dbl, _dbl]
        L.orc_err_class.restype = _i32
        L.orc_err_class.argtypes = [_dbl]
        L.orc_nmax_table.argtypes = [_p, _dbl, _p]
        L.orc_mapped_ratio.restype = _dbl
        L.orc_mapped_ratio.argtypes = [_i64, _i64, _p, _p, _i64, _dbl, _dbl]
        L.orc_total_mapped.restype = _i64
        L.orc_total_mapped.argtypes = [_i64, _i64, _p, _p, _i64, _i32]
        L.orc_set_gap_open.restype = _i32
        L.orc_set_gap_open.argtypes = [_dbl]
        L.orc_align.restype = _i64
        L.orc_align.argtypes = [_p, _i64, _p, _i64, _i32, _dbl, _i32, _p, _p, _p, _p, _p]
        L.orc_aligned_windows.restype = _i64
        L.orc_aligned_windows.argtypes = [_p, _i64, _dbl, _i32]
        L.orc_cluster_reads.restype = _i64
        L.orc_cluster_reads.argtypes = [_p, _p, _p, _p, _p, _p]
        L.orc_mindb_after_updates.restype = _i64
        L.orc_mindb_after_updates.argtypes = [_p, _p, _i32, _p, _p, _p, _i32, _p, _p, _p, _i64, _i64]
        L.orc_cluster_seeded.restype = _i64
        L.orc_cluster_seeded.argtypes = [_p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _p, _p, _p, _p]
        self.tab = np.zeros(129)
        self.tab_nomin = np.zeros(129)
        L.orc_init_qualtab(_ptr(self.tab), 0)
        L.orc_init_qualtab(_ptr(self.tab_nomin), 1)

    def qual_score(self, qual, k):
        q = _bytes(qual)
        return self.L.orc_calc_qual_score(_ptr(q), len(q), k, _ptr(self.tab))

    def error_rate(self, qual, nomin=True):
        q = _bytes(qual)
        return self.L.orc_calc_error_rate(_ptr(q), len(q), _ptr(self.tab_nomin if nomin else self.tab))

    def fill_scores(self, qual, offs, k):
        n = len(offs) - 1
        score, err = np.zeros(n), np.zeros(n)
        s, e = _dbl(), _dbl()
        for i in range(n):
            a, b = int(offs[i]), int(offs[i + 1])
            self.L.orc_fill_qual_score(qual.ctypes.data + a, b - a, k, _ptr(self.tab), _ptr(self.tab_nomin), C.byref(s), C.byref(e))
            score[i], err[i] = s.value, e.value
        return score, err

    def sort_order(self, score):
        score = np.ascontiguousarray(score, dtype=np.float64)
        order = np.zeros(len(score), dtype=np.int32)
        self.L.orc_sort_by_score(_ptr(score), len(score), _ptr(order))
        return order

    def hpc(self, seq, qual):
        s, q = _bytes(seq), _bytes(qual)
        os_, oq = np.zeros(len(s), np.uint8), np.zeros(len(s), np.uint8)
        n = self.L.orc_hpc(_ptr(s), _ptr(q), len(s), _ptr(os_), _ptr(oq))
        return os_[:n].copy(), oq[:n].copy()

    def revcomp(self, seq):
        s = _bytes(seq)
        o = np.zeros(len(s), np.uint8)
        n = self.L.orc_revcomp(_ptr(s), len(s), _ptr(o))
        return None if n < 0 else o

    def kmer_encode(self, seq, k):
        s = _bytes(seq)
        o = np.zeros(max(len(s), 1), np.uint32)
        n = self.L.orc_kmer_encode(_ptr(s), len(s), k, _ptr(o))
        return o[:n].copy()

    def minimizers(self, kmers, k, w):
        kv = np.ascontiguousarray(kmers, dtype=np.uint32)
        o = np.zeros(max(len(kv), 1), MIN_DT)
        n = self.L.orc_minimizers(_ptr(kv), len(kv), k, w, _ptr(o))
        return o[:n].copy()

    def sketch_read(self, seq, qual, k, w):
        s, q = _bytes(seq), _bytes(qual)
        n = len(s)
        hs, hq = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.uint8)
        mins, rmins = np.zeros(max(n, 1), MIN_DT), np.zeros(max(n, 1), MIN_DT)
        hl, nm, nr = _i64(), _i64(), _i64()
        he = _dbl()
        st = self.L.orc_sketch_read(_ptr(s), _ptr(q), n, k, w, _ptr(self.tab_nomin), _ptr(hs), _ptr(hq), C.byref(hl),
                                    C.byref(he), _ptr(mins), C.byref(nm), _ptr(rmins), C.byref(nr))
        return dict(status=st, hpc_seq=hs[:hl.value].copy(), hpc_qual=hq[:hl.value].copy(), hpc_err=he.value,
                    mins=mins[:nm.value].copy(), rev_mins=rmins[:nr.value].copy())

    def sketch_batch(self, seq, qual, offs, k, w):
        """Sketch every read (no quality filter); returns flat arrays like the product's download."""
        n = len(offs) - 1
        status = np.zeros(n, np.int32)
        hpc_err = np.full(n, -1.0)
        hs, hq, ms, rs = [], [], [], []
        hoffs, moffs, roffs = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
        for i in range(n):
            a, b = int(offs[i]), int(offs[i + 1])
            r = self.sketch_read(seq[a:b], qual[a:b], k, w)
            status[i] = r["status"]
            if r["status"] == 1:
                hpc_err[i] = r["hpc_err"]
                hs.append(r["hpc_seq"]); hq.append(r["hpc_qual"]); ms.append(r["mins"]); rs.append(r["rev_mins"])
                hoffs[i + 1] = hoffs[i] + len(r["hpc_seq"])
                moffs[i + 1] = moffs[i] + len(r["mins"])
                roffs[i + 1] = roffs[i] + len(r["rev_mins"])
            else:
                hoffs[i + 1], moffs[i + 1], roffs[i + 1] = hoffs[i], moffs[i], roffs[i]
        cat = lambda l, dt: np.concatenate(l) if l else np.zeros(0, dt)
        return dict(status=status, hpc_err=hpc_err, hpc_seq=cat(hs, np.uint8), hpc_qual=cat(hq, np.uint8), hpc_offs=hoffs,
                    mins=cat(ms, MIN_DT), min_offs=moffs, rev_mins=cat(rs, MIN_DT), rev_offs=roffs)

    def nmax_table(self, p_shared, min_prob_no_hits):
        p = np.ascontiguousarray(p_shared, dtype=np.float64).reshape(225)
        o = np.zeros(225, np.int32)
        self.L.orc_nmax_table(_ptr(p), min_prob_no_hits, _ptr(o))
        return o.reshape(15, 15)

    def mapped_ratio(self, hpc_len, n_mins, pos, idx, p_shared, min_prob_no_hits):
        pos = np.ascontiguousarray(pos, np.uint32); idx = np.ascontiguousarray(idx, np.uint32)
        return self.L.orc_mapped_ratio(hpc_len, n_mins, _ptr(pos), _ptr(idx), len(pos), p_shared, min_prob_no_hits)

    def total_mapped(self, hpc_len, n_mins, pos, idx, nmax):
        pos = np.ascontiguousarray(pos, np.uint32); idx = np.ascontiguousarray(idx, np.uint32)
        return self.L.orc_total_mapped(hpc_len, n_mins, _ptr(pos), _ptr(idx), len(pos), int(nmax))

    def align(self, read, ref, gap_open, e, k, want_comp=False):
        a, b = _bytes(read), _bytes(ref)
        score = C.c_int32(); eq = _i64(); er = _i64(); cl = _i64()
        comp = np.zeros(len(a) + len(b) + 1, np.uint8) if want_comp else None
        al = self.L.orc_align(_ptr(a), len(a), _ptr(b), len(b), gap_open, e, k, C.byref(score), C.byref(eq), C.byref(er),
                              _ptr(comp), C.byref(cl))
        return dict(aligned=al, score=score.value, end_query=eq.value, end_ref=er.value, comp_len=cl.value,
                    comp=comp[:cl.value].tobytes() if want_comp else None)

    def mindb_after_updates(self, cls_mins, cls_offs, upd_cls, upd_mins, upd_offs):
        return _mindb_after_updates(self.L.orc_mindb_after_updates, cls_mins, cls_offs, upd_cls, upd_mins, upd_offs)

    def _params(self, k, w, mode, p_shared, min_shared, mapped_th, aligned_th, min_frac, min_prob_no_hits, min_qual):
        P = ClusterParams(k=k, w=w, mode=mode, min_shared=min_shared, mapped_th=mapped_th, aligned_th=aligned_th,
                          min_frac=min_frac, min_prob_no_hits=min_prob_no_hits, min_qual=min_qual)
        ps = np.ascontiguousarray(p_shared, np.float64).reshape(225)
        for i in range(225):
            P.p_shared[i] = ps[i]
        return P

    def _reads(self, seq, offs, sk, score, raw_err, order, status):
        """orc_reads over the reads `order` (indices into the flat read set), in that order; returns (struct, keepalive)."""
        n = len(order)
        order = np.asarray(order)
        lens = (offs[1:] - offs[:-1])[order]
        raw_offs = np.zeros(n + 1, np.int64); raw_offs[1:] = np.cumsum(lens)
        raw = np.concatenate([seq[offs[i]:offs[i + 1]] for i in order]) if n else np.zeros(0, np.uint8)
        hl = (sk["hpc_offs"][1:] - sk["hpc_offs"][:-1])[order].astype(np.int64)
        def gather(arr, o):
            cnt = (o[1:] - o[:-1])[order]
            no = np.zeros(n + 1, np.int64); no[1:] = np.cumsum(cnt)
            parts = [arr[o[i]:o[i + 1]] for i in order]
            return (np.concatenate(parts) if parts else np.zeros(0, MIN_DT)), no
        mins, moffs = gather(sk["mins"], sk["min_offs"])
        rmins, roffs = gather(sk["rev_mins"], sk["rev_offs"])
        st = np.ascontiguousarray((sk["status"] if status is None else status)[order], np.int32)
        sc = np.ascontiguousarray(np.asarray(score)[order], np.float64)
        re_ = np.ascontiguousarray(np.asarray(raw_err)[order], np.float64)
        he = np.ascontiguousarray(sk["hpc_err"][order], np.float64)
        mins = np.ascontiguousarray(mins); rmins = np.ascontiguousarray(rmins); raw = np.ascontiguousarray(raw)
        R = Reads(n_reads=n, status=_ptr(st), score=_ptr(sc), raw_err=_ptr(re_), hpc_err=_ptr(he), raw_seq=_ptr(raw),
                  raw_offs=_ptr(raw_offs), hpc_len=_ptr(hl), mins=_ptr(mins), min_offs=_ptr(moffs), rev_mins=_ptr(rmins),
                  rev_offs=_ptr(roffs))
        return R, (st, sc, re_, he, raw, raw_offs, hl, mins, moffs, rmins, roffs)

    def cluster(self, seq, offs, sk, score, raw_err, order, k, w, mode, p_shared, min_shared=5, mapped_th=0.65,
                aligned_th=0.2, min_frac=0.8, min_prob_no_hits=0.1, min_qual=7.0, status=None):
        """Greedy clustering of reads given in SORTED order `order` (indices into the flat read set).

        sk: dict from sketch_batch over the ORIGINAL read order.  Returns dict with per-sorted-slot results."""
        n = len(order)
        P = self._params(k, w, mode, p_shared, min_shared, mapped_th, aligned_th, min_frac, min_prob_no_hits, min_qual)
        R, _keep = self._reads(seq, offs, sk, score, raw_err, order, status)
        oc, os_, of = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        na = _i64()
        nc = self.L.orc_cluster_reads(C.byref(P), C.byref(R), _ptr(oc), _ptr(os_), _ptr(of), C.byref(na))
        return dict(n_clusters=nc, cluster=oc, strand=os_, founder=of[:nc].copy(), n_aligned=na.value)

    def cluster_seeded(self, seq, offs, sk, score, raw_err, seeds, updates, queries, k, w, mode, p_shared, min_shared=5,
                       mapped_th=0.65, aligned_th=0.2, min_frac=0.8, min_prob_no_hits=0.1, min_qual=7.0, status=None):
        """Clusters seeded from the reads `seeds`, representatives replaced by `updates` [(cluster, read)] (UpdateMinDB), then the
        reads `queries` clustered in that order.  All read ids index the flat read set.  Returns per-query results."""
        P = self._params(k, w, mode, p_shared, min_shared, mapped_th, aligned_th, min_frac, min_prob_no_hits, min_qual)
        R, _keep = self._reads(seq, offs, sk, score, raw_err, np.arange(len(offs) - 1), status)
        seeds = np.ascontiguousarray(seeds, np.int32)
        uc = np.ascontiguousarray([c for c, _ in updates], np.int32)
        ur = np.ascontiguousarray([r for _, r in updates], np.int32)
        q = np.ascontiguousarray(queries, np.int32)
        oc, os_ = np.zeros(len(q), np.int32), np.zeros(len(q), np.int32)
        of = np.zeros(len(seeds) + len(q), np.int32)
        na = _i64()
        nc = self.L.orc_cluster_seeded(C.byref(P), C.byref(R), _ptr(seeds), len(seeds), _ptr(uc), _ptr(ur), len(uc), _ptr(q), len(q),
                                       _ptr(oc), _ptr(os_), _ptr(of), C.byref(na))
        assert nc >= 0
        return dict(n_clusters=nc, cluster=oc, strand=os_, founder=of[:nc].copy(), n_aligned=na.value)


class Ref:
    """The unmodified reference (oracle/_ref)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise RuntimeError(f"{REF_SO} missing (built only where /root/reference exists)")
        L = self.L = C.CDLL(REF_SO)
        L.ref_hpc.restype = _i32
        L.ref_hpc.argtypes = [_p, _p, _i32, _p, _p]
        L.ref_revcomp.restype = _i32
        L.ref_revcomp.argtypes = [_p, _i32, _p]
        L.ref_kmer_encode.restype = _i32
        L.ref_kmer_encode.argtypes = [_p, _i32, _i32, _p]
        L.ref_minimizers.restype = _i32
        L.ref_minimizers.argtypes = [_p, _i32, _i32, _i32, _p]
        L.ref_qualtabs.argtypes = [_p, _p]
        L.ref_calc_qual_score.restype = _dbl
        L.ref_calc_qual_score.argtypes = [_p, _i32, _i32]
        L.ref_calc_error_rate.restype = _dbl
        L.ref_calc_error_rate.argtypes = [_p, _i32, _i32]
        L.ref_pmin_shared.restype = _dbl
        L.ref_pmin_shared.argtypes = [_i32, _i32, _dbl, _dbl]
        L.ref_mapped_ratio.restype = _dbl
        L.ref_mapped_ratio.argtypes = [_i32, _i32, _i32, _dbl, _dbl, _i32, _p, _i32, _dbl]
        L.ref_set_gap_open.restype = _i32
        L.ref_set_gap_open.argtypes = [_dbl]
        L.ref_align.restype = _i32
        L.ref_align.argtypes = [_p, _i32, _p, _i32, _i32, _dbl, _i32, _p, _p, _p, _p, _p, _p, _p]
        L.ref_align_scalar_trace.restype = _i32
        L.ref_align_scalar_trace.argtypes = [_p, _i32, _p, _i32, _i32, _p, _p, _p]
        L.ref_sort_and_sketch.restype = _p
        L.ref_sort_and_sketch.argtypes = [_p, _p, _p, _i32, _i32, _i32, _dbl, _i32]
        L.ref_batch_free.argtypes = [_p]
        L.ref_batch_scores.argtypes = [_p, _p, _p, _p]
        L.ref_batch_sketch_sizes.argtypes = [_p, _p, _p, _p]
        L.ref_batch_sketch_export.argtypes = [_p] * 11
        L.ref_batch_slice.restype = _p
        L.ref_batch_slice.argtypes = [_p, _i32, _i32, _i32]
        L.ref_cluster.restype = _i32
        L.ref_cluster.argtypes = [_p, _p, _i32, _i32, _dbl, _dbl, _dbl, _dbl, _i32, _i32]
        L.ref_aln_invoked.restype = C.c_uint
        L.ref_n_clusters.restype = _i32
        L.ref_n_clusters.argtypes = [_p]
        L.ref_cluster_sizes.argtypes = [_p, _p]
        L.ref_cluster_members.argtypes = [_p, _p, _p]
        L.ref_mindb_size.restype = _i64
        L.ref_mindb_size.argtypes = [_p, _p]
        L.ref_mindb_export.argtypes = [_p, _p, _p, _p]
        L.ref_save_batch.argtypes = [_p, C.c_char_p]
        L.ref_load_batch.restype = _p
        L.ref_load_batch.argtypes = [C.c_char_p]
        L.ref_sorted_hits.restype = _i32
        L.ref_sorted_hits.argtypes = [_p, _p, _i32, _p, _i32, _p, _i32, _p]
        L.ref_mindb_after_updates.restype = _i64
        L.ref_mindb_after_updates.argtypes = [_p, _p, _i32, _p, _p, _p, _i32, _p, _p, _p, _i64, _i64]

    def qualtabs(self):
        a, b = np.zeros(129), np.zeros(129)
        self.L.ref_qualtabs(_ptr(a), _ptr(b))
        return a, b

    def hpc(self, seq, qual):
        s, q = _bytes(seq), _bytes(qual)
        os_, oq = np.zeros(len(s), np.uint8), np.zeros(len(s), np.uint8)
        n = self.L.ref_hpc(_ptr(s), _ptr(q), len(s), _ptr(os_), _ptr(oq))
        return os_[:n].copy(), oq[:n].copy()

    def revcomp(self, seq):
        s = _bytes(seq)
        o = np.zeros(len(s), np.uint8)
        n = self.L.ref_revcomp(_ptr(s), len(s), _ptr(o))
        return None if n < 0 else o

    def kmer_encode(self, seq, k):
        s = _bytes(seq)
        o = np.zeros(max(len(s), 1), np.uint32)
        n = self.L.ref_kmer_encode(_ptr(s), len(s), k, _ptr(o))
        return o[:n].copy()

    def minimizers(self, kmers, k, w):
        kv = np.ascontiguousarray(kmers, dtype=np.uint32)
        o = np.zeros(max(len(kv), 1), MIN_DT)
        n = self.L.ref_minimizers(_ptr(kv), len(kv), k, w, _ptr(o))
        return o[:n].copy()

    def qual_score(self, qual, k):
        q = _bytes(qual)
        return self.L.ref_calc_qual_score(_ptr(q), len(q), k)

    def error_rate(self, qual, nomin=True):
        q = _bytes(qual)
        return self.L.ref_calc_error_rate(_ptr(q), len(q), 1 if nomin else 0)

    def pmin_shared(self, k, w, e1, e2):
        return self.L.ref_pmin_shared(k, w, e1, e2)

    def mapped_ratio(self, k, w, hpc_len, read_err, cls_err, n_mins, pos, idx, min_prob_no_hits):
        h = np.ascontiguousarray(np.stack([pos, idx], axis=1), np.uint32)
        return self.L.ref_mapped_ratio(k, w, hpc_len, read_err, cls_err, n_mins, _ptr(h), len(h), min_prob_no_hits)

    def set_gap_open(self, e):
        return self.L.ref_set_gap_open(e)

    def align(self, read, ref, gap_open, e, k, want_comp=False):
        a, b = _bytes(read), _bytes(ref)
        score, eq, er, cl, sat = _i32(), _i32(), _i32(), _i32(), _i32()
        ratio = _dbl()
        comp = np.zeros(len(a) + len(b) + 1, np.uint8) if want_comp else None
        al = self.L.ref_align(_ptr(a), len(a), _ptr(b), len(b), gap_open, e, k, C.byref(score), C.byref(eq),
                              C.byref(er), _ptr(comp), C.byref(cl), C.byref(ratio), C.byref(sat))
        return dict(aligned=al, score=score.value, end_query=eq.value, end_ref=er.value, comp_len=cl.value,
                    ratio=ratio.value, saturated16=sat.value, comp=comp[:cl.value].tobytes() if want_comp else None)

    def align_scalar_trace(self, read, ref, gap_open):
        a, b = _bytes(read), _bytes(ref)
        tr = np.zeros((len(a), len(b)), np.int8)
        eq, er = _i32(), _i32()
        sc = self.L.ref_align_scalar_trace(_ptr(a), len(a), _ptr(b), len(b), gap_open, _ptr(tr), C.byref(eq), C.byref(er))
        return sc, eq.value, er.value, tr

    # ---- batches
    def sort_and_sketch(self, seq, qual, offs, k, w, min_qual=7.0, sketch=True):
        seq, qual = _bytes(seq), _bytes(qual)
        offs = np.ascontiguousarray(offs, np.int64)
        return RefBatch(self, self.L.ref_sort_and_sketch(_ptr(seq), _ptr(qual), _ptr(offs), len(offs) - 1, k, w, min_qual,
                                                          1 if sketch else 0), len(offs) - 1)

    def mindb_after_updates(self, cls_mins, cls_offs, upd_cls, upd_mins, upd_offs):
        return _mindb_after_updates(self.L.ref_mindb_after_updates, cls_mins, cls_offs, upd_cls, upd_mins, upd_offs)

    def load_batch(self, path):
        """LoadBatch (serialize.cpp:12-27) of a `.cer` written by RefBatch.save or by either CLI."""
        return RefBatch(self, self.L.ref_load_batch(path.encode()), 0)

    def sorted_hits(self, cls_mins, cls_offs, mins, rev_mins):
        cm = np.ascontiguousarray(cls_mins); co = np.ascontiguousarray(cls_offs, np.int64)
        a = np.ascontiguousarray(mins); b = np.ascontiguousarray(rev_mins)
        out = np.zeros((2 * (len(co) - 1) + 1, 3), np.int32)
        n = self.L.ref_sorted_hits(_ptr(cm), _ptr(co), len(co) - 1, _ptr(a), len(a), _ptr(b), len(b), _ptr(out))
        return out[:n].copy()


class RefBatch:
    def __init__(self, ref, handle, n):
        self.ref, self.h, self.n = ref, handle, n

    def __del__(self):
        if self.h:
            self.ref.L.ref_batch_free(self.h)
            self.h = None

    def scores(self):
        s, e, o = np.zeros(self.n), np.zeros(self.n), np.zeros(self.n, np.int32)
        self.ref.L.ref_batch_scores(self.h, _ptr(s), _ptr(e), _ptr(o))
        return s, e, o

    def sketch(self):
        """Flat export in SORTED slot order."""
        a, b, c = _i64(), _i64(), _i64()
        self.ref.L.ref_batch_sketch_sizes(self.h, C.byref(a), C.byref(b), C.byref(c))
        n = self.n
        status = np.zeros(n, np.int32)
        hs, hq = np.zeros(max(a.value, 1), np.uint8), np.zeros(max(a.value, 1), np.uint8)
        ho, mo, ro = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
        he, rs = np.zeros(n), np.zeros(n)
        mins, rmins = np.zeros(max(b.value, 1), MIN_DT), np.zeros(max(c.value, 1), MIN_DT)
        self.ref.L.ref_batch_sketch_export(self.h, _ptr(status), _ptr(hs), _ptr(hq), _ptr(ho), _ptr(he), _ptr(rs), _ptr(mins),
                                           _ptr(mo), _ptr(rmins), _ptr(ro))
        return dict(status=status, hpc_seq=hs[:a.value], hpc_qual=hq[:a.value], hpc_offs=ho, hpc_err=he, raw_score=rs,
                    mins=mins[:b.value], min_offs=mo, rev_mins=rmins[:c.value], rev_offs=ro)

    def slice(self, lo, hi, batch_nr):
        return RefBatch(self.ref, self.ref.L.ref_batch_slice(self.h, lo, hi, batch_nr), hi - lo + 1)

    def cluster(self, mode, right=None, min_shared=5, mapped_th=0.65, aligned_th=0.2, min_frac=0.8, min_prob_no_hits=0.1,
                min_cls_size=-1, spoa_algo=2):
        before = self.ref.L.ref_aln_invoked()
        nc = self.ref.L.ref_cluster(self.h, right.h if right is not None else None, mode, min_shared, mapped_th,
                                    aligned_th, min_frac, min_prob_no_hits, min_cls_size, spoa_algo)
        sizes = np.zeros(nc, np.int32)
        self.ref.L.ref_cluster_sizes(self.h, _ptr(sizes))
        tot = int(sizes.sum())
        ids, strands = np.zeros(tot, np.int32), np.zeros(tot, np.int32)
        self.ref.L.ref_cluster_members(self.h, _ptr(ids), _ptr(strands))
        return dict(n_clusters=nc, sizes=sizes, ids=ids, strands=strands,
                    n_aligned=self.ref.L.ref_aln_invoked() - before)

    def cluster_lists(self):
        """The cluster lists of a batch as it stands (e.g. loaded from a `.cer`): sizes, member read ids, strands."""
        nc = self.ref.L.ref_n_clusters(self.h)
        sizes = np.zeros(nc, np.int32)
        self.ref.L.ref_cluster_sizes(self.h, _ptr(sizes))
        tot = int(sizes.sum())
        ids, strands = np.zeros(tot, np.int32), np.zeros(tot, np.int32)
        self.ref.L.ref_cluster_members(self.h, _ptr(ids), _ptr(strands))
        return dict(n_clusters=nc, sizes=sizes, ids=ids, strands=strands)

    def mindb(self):
        t = _i64()
        n = self.ref.L.ref_mindb_size(self.h, C.byref(t))
        keys, offs, cls = np.zeros(n, np.uint32), np.zeros(n + 1, np.int64), np.zeros(max(t.value, 1), np.uint32)
        self.ref.L.ref_mindb_export(self.h, _ptr(keys), _ptr(offs), _ptr(cls))
        return keys, offs, cls[:t.value]

    def save(self, path):
        self.ref.L.ref_save_batch(self.h, path.encode())


_port = _ref = None


def port():
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref():
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def have_ref():
    return os.path.exists(REF_SO)
