"""GPU: the C++ header face (include/isonclust2_compat/) and the per-primitive entry points of the C ABI behind it.

test_reference_unit_test runs the reference's OWN unit test (test/isONclust2_test.cpp, unmodified, compiled in the dev
container against the shim, oracle/_ref/compat_unit_test): its 8 tests pin sorting by score, minimizers, homopolymer
compression, error rates, the empirical probability lookup, minimizer matching + mapped ratio, the alignment ratio and the
k-mer transform on the reference's golden values.  The other tests compare every primitive with the oracle on seeded inputs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from isonclust2_b200 import api, pemp, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HIT_DT = np.dtype([("cls", "<u4"), ("strand", "<i4"), ("pos", "<u4"), ("index", "<u4")])


def test_reference_unit_test():
    exe = os.path.join(ROOT, "oracle", "_ref", "compat_unit_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/compat_unit_test not built (needs /root/reference at build time)")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "8 tests, 0 failed" in r.stdout, r.stdout[-3000:]


def _lib(ctx):
    L = ctx.L
    return L


def test_primitives_match_oracle(port, ctx):
    L = ctx.L
    k, w = 13, 20
    seq, qual, offs = synth.make_reads(40, seed=301, min_len=200, max_len=900)
    n = len(offs) - 1
    # HomopolymerCompress
    hs, hq, hl = np.zeros(len(seq), np.uint8), np.zeros(len(seq), np.uint8), np.zeros(n, np.int32)
    ctx.check(L.icb_hpc_compress(ctx.h, seq.ctypes.data, qual.ctypes.data, offs.ctypes.data, n, hs.ctypes.data, hq.ctypes.data, hl.ctypes.data))
    tab = port.tab
    for i in range(n):
        a, b = int(offs[i]), int(offs[i + 1])
        ps, pq = port.hpc(seq[a:b], qual[a:b])
        assert hl[i] == len(ps) and (hs[a:a + hl[i]] == ps).all() and (hq[a:a + hl[i]] == pq).all()
        # KmerEncodeSeq + GetKmerMinimizers on the compressed read
        km = np.zeros(max(len(ps), 1), np.uint32)
        nk = C.c_int64()
        ctx.check(L.icb_kmer_encode(ctx.h, ps.ctypes.data, len(ps), k, km.ctypes.data, C.byref(nk)))
        pk = port.kmer_encode(ps, k)
        assert nk.value == len(pk) and (km[:nk.value] == pk).all()
        mins = np.zeros(max(nk.value, 1), api.MIN_DT)
        nm = C.c_int64()
        ctx.check(L.icb_kmer_minimizers(ctx.h, km.ctypes.data, nk.value, k, w, mins.ctypes.data, C.byref(nm)))
        pm = port.minimizers(pk, k, w)
        assert nm.value == len(pm) and (mins[:nm.value] == pm).all()
    # CalcErrorRate of every quality string, capped and uncapped table
    for t in (port.tab, port.tab_nomin):
        out = np.zeros(n)
        ctx.check(L.icb_calc_error_rate(ctx.h, qual.ctypes.data, offs.ctypes.data, n, t.ctypes.data, out.ctypes.data))
        want = np.array([port.error_rate(qual[offs[i]:offs[i + 1]], nomin=t is port.tab_nomin) for i in range(n)])
        assert (out == want).all()
    assert tab[33] == 0.79433


def test_index_primitives_match_oracle(port, ctx):
    """AddMinimizers / UpdateMinDB / GetMinimizerHits through the device index against the oracle's MinDB, then totalMapped."""
    L = ctx.L
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(60, seed=302, reads_per_gene=6, min_len=300, max_len=700)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    good = [i for i in range(len(offs) - 1) if sk["status"][i] == 1]
    reps, queries = good[:12], good[12:30]
    sketch = lambda i: np.ascontiguousarray(sk["mins"][sk["min_offs"][i]:sk["min_offs"][i + 1]])
    rsketch = lambda i: np.ascontiguousarray(sk["rev_mins"][sk["rev_offs"][i]:sk["rev_offs"][i + 1]])
    ix = C.c_void_p()
    ctx.check(L.icb_index_create(ctx.h, C.byref(ix)))
    try:
        for c, r in enumerate(reps):
            m = sketch(r)
            ctx.check(L.icb_index_add(ix, c, m.ctypes.data, len(m)))
        upd = [(3, good[31]), (7, good[32]), (3, good[33])]
        cur = list(reps)
        for c, r in upd:
            o, nw = sketch(cur[c]), sketch(r)
            ctx.check(L.icb_index_update_cluster(ix, c, o.ctypes.data, len(o), nw.ctypes.data, len(nw)))
            cur[c] = r
        # the oracle's MinDB after the same operations
        cm = np.concatenate([sketch(r) for r in reps])
        co = np.zeros(len(reps) + 1, np.int64)
        co[1:] = np.cumsum([len(sketch(r)) for r in reps])
        um = np.concatenate([sketch(r) for _, r in upd])
        uo = np.zeros(len(upd) + 1, np.int64)
        uo[1:] = np.cumsum([len(sketch(r)) for _, r in upd])
        keys, koffs, kcls = port.mindb_after_updates(cm, co, [c for c, _ in upd], um, uo)
        db = {int(keys[i]): kcls[koffs[i]:koffs[i + 1]].tolist() for i in range(len(keys))}
        nmax = port.nmax_table(pemp.load_p_shared(k, w), 0.1)
        for q in queries:
            f, r = sketch(q), rsketch(q)
            hits = np.zeros(64, HIT_DT)
            nh = C.c_int64()
            ctx.check(L.icb_index_query_hits(ix, f.ctypes.data, len(f), r.ctypes.data, len(r), hits.ctypes.data, len(hits), C.byref(nh)))
            if nh.value > len(hits):
                hits = np.zeros(nh.value, HIT_DT)
                ctx.check(L.icb_index_query_hits(ix, f.ctypes.data, len(f), r.ctypes.data, len(r), hits.ctypes.data, len(hits), C.byref(nh)))
            hits = hits[:nh.value]
            want = []
            for strand, arr in ((1, f), (-1, r)):
                for m in arr:
                    for c in db.get(int(m["Min"]), []):
                        want.append((c, strand, int(m["Pos"]), int(m["Index"])))
            assert [tuple(int(x) for x in h) for h in hits] == want
            # totalMapped of the best forward candidate
            fw = hits[hits["strand"] == 1]
            if len(fw):
                c = np.bincount(fw["cls"]).argmax()
                hv = fw[fw["cls"] == c]
                hpc_len = int(sk["hpc_offs"][q + 1] - sk["hpc_offs"][q])
                cls_ = port.L.orc_err_class(sk["hpc_err"][cur[c]]) * 15 + port.L.orc_err_class(sk["hpc_err"][q])
                nm = int(nmax.reshape(225)[cls_])
                tot = C.c_int64()
                pos, idx = np.ascontiguousarray(hv["pos"]), np.ascontiguousarray(hv["index"])
                ctx.check(L.icb_mapped_length(ctx.h, pos.ctypes.data, idx.ctypes.data, len(hv), len(f), hpc_len, nm, C.byref(tot)))
                assert tot.value == port.total_mapped(hpc_len, len(f), pos, idx, nm)
    finally:
        L.icb_index_destroy(ix)


def test_align_traceback_comp_matches_oracle(port, ctx):
    """icb_align_traceback: score, end cell and the materialised comparison string against the oracle's traceback, and
    icb_count_windows on that string against the oracle's window count."""
    L = ctx.L
    k = 11
    seq, qual, offs = synth.make_reads(16, seed=303, reads_per_gene=4, min_len=150, max_len=600)
    rs = ctx.upload(seq, qual, offs)
    try:
        for a, b, strand in ((0, 1, 1), (2, 3, 1), (4, 9, 1), (5, 6, -1), (7, 7, 1)):
            read = seq[offs[a]:offs[a + 1]]
            rep = seq[offs[b]:offs[b + 1]]
            if strand == -1:
                rep = port.revcomp(rep)
            e = 0.11
            go = api.set_gap_open(e)
            want = port.align(read, rep, go, e, k, want_comp=True)
            pair = np.array([(a, b, strand, go, api.aln_limit(e, k))], dtype=api.PAIR_DT)
            res = np.zeros(1, api.ALN_DT)
            comp = np.zeros(len(read) + len(rep) + 1, np.uint8)
            ctx.check(L.icb_align_traceback(rs.h, pair.ctypes.data, k, res.ctypes.data, comp.ctypes.data, len(comp)))
            assert int(res["score"][0]) == want["score"] and int(res["comp_len"][0]) == want["comp_len"]
            assert int(res["end_query"][0]) == want["end_query"] and int(res["end_ref"][0]) == want["end_ref"]
            assert comp[:want["comp_len"]].tobytes() == want["comp"]
            assert int(res["aligned"][0]) == want["aligned"]
            al = C.c_int64()
            ctx.check(L.icb_count_windows(ctx.h, comp.ctypes.data, want["comp_len"], k, api.aln_limit(e, k), C.byref(al)))
            assert al.value == want["aligned"]
    finally:
        rs.close()


def test_load_p_shared_matches_python_loader(ctx):
    for k, w in ((11, 15), (13, 20), (10, 14)):
        g = np.zeros(225)
        assert ctx.L.icb_load_p_shared(None, k, w, g.ctypes.data) == 0
        want = pemp.load_p_shared(k, w).reshape(225)
        assert ((g == want) | (np.isnan(g) & np.isnan(want))).all()
