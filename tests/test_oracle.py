"""CPU suite: pins the oracle restatement (oracle/isonclust2_oracle.c[pp]) against
  (1) the reference's own unit-test goldens (test/isONclust2_test.cpp:17-203),
  (2) fixtures produced by the unmodified reference (tests/golden/make_golden.py),
  (3) the unmodified reference itself (oracle/_ref) on seeded random inputs, when it is built here.
"""
import numpy as np
import pytest

from isonclust2_b200 import pemp, synth
from oracle import MIN_DT

REF = ("GGTAGTGGTGGCGGGTCTCCTTGAGAGCACTCGTCGAGTATGCCGAAAATATGTTAATGGCAGGAAGTTTGATTATAGCCATTAGCGTGTCATAATGTAGAAAGTCTCGATAATAAAGCT"
       "CAGGACGCGCCTCCGTTAAAGGAAGGCGGGATCCTGCGCGATGGCTATCTATAGTATGTAGTTACCTCTGATTGTCATGTGAACAGGAGGCCAGTACCACCTGATACGGCCTTGTAAACC"
       "TACCACTACTTCGCTTAAGACGGTGCTCCCCTCCCCATTTGCGGCCGTTCGTCGTGTCCC").encode()
READ = ("AGATATTATAGCCATACGTGTCATAATGTAGAAGTCTCGATAATAAAGCTCAGGACGCGCCTCCGTTAAGGAAGGCGGATCCGCGCGATGGGCTATCTATAGTATGTGGTTACCCTGATA"
        "GTCATGTGAGACAGGAGGCCAGTCCACCTGATACGGCTTGTAAACTACCACTACTTCGCT").encode()


def ulp_eq(a, b, ulps=4):
    ia, ib = np.float64(a).view(np.int64), np.float64(b).view(np.int64)
    return abs(int(ia) - int(ib)) <= ulps


# ---------------------------------------------------------------- (1) the reference's unit-test goldens
def test_sorting_golden(port):  # SortingTest :17-43
    reads = [(b"ATGCGCATATGCGC", b"@IIIIIIIIIIIII"), (b"ATGCTGACATGCATGC", b"@IIIIIIIIIIIIIII"),
             (b"ATGCATGCCGATGTACATGCATGCATCGACGT", b"I" * 32)]
    qual = np.frombuffer(b"".join(q for _, q in reads), np.uint8)
    offs = np.cumsum([0] + [len(s) for s, _ in reads]).astype(np.int64)
    score, _ = port.fill_scores(qual, offs, 3)
    assert list(port.sort_order(score)) == [2, 1, 0]


def test_minimizer_golden(port):  # MinimizerTest :46-57
    m = port.minimizers(port.kmer_encode(b"ACGCCGATC", 2), 2, 4)
    assert m.tolist() == [(1, 0, 0), (5, 3, 1), (3, 6, 2)]  # AC=1, CC=5, AT=3


def test_hpc_golden(port):  # HpcTest :60-67
    s, q = port.hpc(b"AAATTTGCGTTAA", b"++:+?++++++@+")
    assert s.tobytes() == b"ATGCGTA" and q.tobytes() == b":?++++@"


def test_error_rate_golden(port):  # ErrorRateTest :70-74
    assert ulp_eq(port.error_rate(b"I" * 18, nomin=False), 0.0001)


def test_emp_prob_golden():  # EmpProbLookupTest :77-82
    t = pemp.load_p_shared(13, 20)
    assert t[round(0.111 * 100) - 1, round(0.131 * 100) - 1] == 0.11736487779693013


def test_min_match_golden(port):  # MinMatchTest :85-135
    k, w = 13, 20
    rh, _ = port.hpc(REF, b"I" * len(REF))
    qh, qq = port.hpc(READ, b"I" * len(READ))
    rm = port.minimizers(port.kmer_encode(rh, k), k, w)
    qm = port.minimizers(port.kmer_encode(qh, k), k, w)
    refset = set(rm["Min"].tolist())
    hits = qm[np.isin(qm["Min"], list(refset))]
    assert len(hits) == 14
    e = port.error_rate(qq, nomin=False)
    p = pemp.load_p_shared(k, w)[port.L.orc_err_class(e), port.L.orc_err_class(e)]
    assert ulp_eq(1.0 - p, 0.17140336964776648)
    mr = port.mapped_ratio(len(qh), len(qm), hits["Pos"], hits["Index"], p, 0.1)
    assert ulp_eq(mr, 0.3835616438356164)
    nmax = port.nmax_table(pemp.load_p_shared(k, w), 0.1)[port.L.orc_err_class(e), port.L.orc_err_class(e)]
    assert port.total_mapped(len(qh), len(qm), hits["Pos"], hits["Index"], nmax) / len(qh) == mr


def test_aln_ratio_golden(port):  # AlnRatioTest :137-181
    e1 = port.error_rate(b"I" * len(REF), nomin=False)
    e2 = port.error_rate(b"I" * len(READ), nomin=False)
    go = port.L.orc_set_gap_open(e1 + e2)
    r = port.align(REF, READ, go, e1 + e2, 13)
    assert ulp_eq(r["aligned"] / len(READ), 0.7111111111111111)


def test_kmer_transform_golden(port):  # TestKmerTransform :184-203
    k = 4
    kmers = []
    for i in range(4 ** k):
        s, x = "", i
        for _ in range(k):
            s = "ACGT"[x % 4] + s
            x //= 4
        kmers.append(s)
    assert kmers == sorted(kmers)
    for i, s in enumerate(kmers):
        assert port.kmer_encode((s + "A").encode(), k)[0] == i  # encode drops the last k-mer, so pad one base


# ---------------------------------------------------------------- (2) fixtures made by the unmodified reference
def test_golden_tables(port, golden):
    assert (port.tab == golden["qualtab"]).all() and (port.tab_nomin == golden["qualtab_nomin"]).all()
    assert golden["ut_hpc_seq"].tobytes() == b"ATGCGTA" and golden["ut_hpc_qual"].tobytes() == b":?++++@"
    assert golden["ut_mins"].tolist() == [(1, 0, 0), (5, 3, 1), (3, 6, 2)]
    assert float(golden["ut_pmin"]) == 0.11736487779693013


def test_golden_scores_and_sketch(port, golden):
    seq, qual, offs = golden["seq"], golden["qual"], golden["offs"]
    k, w = int(golden["k"]), int(golden["w"])
    score, err = port.fill_scores(qual, offs, k)
    assert (score == golden["score"]).all() and (err == golden["err"]).all()
    order = port.sort_order(score)
    assert (order == golden["order"]).all()
    sk = port.sketch_batch(seq, qual, offs, k, w)
    filt = np.array([port.L.orc_quality_filtered(e, 7.0) for e in err])
    for slot, i in enumerate(order):
        gs = int(golden["sk_status"][slot])
        if filt[i]:
            assert gs == 0
            continue
        st = int(sk["status"][i])
        assert (st, gs) in ((1, 1), (2, 2), (3, 0)), (i, st, gs)  # status 3 = reference drops raw+hpc
        if st != 1:
            continue
        a, b = golden["sk_min_offs"][slot], golden["sk_min_offs"][slot + 1]
        assert (golden["sk_mins"][a:b] == sk["mins"][sk["min_offs"][i]:sk["min_offs"][i + 1]]).all()
        a, b = golden["sk_rev_offs"][slot], golden["sk_rev_offs"][slot + 1]
        assert (golden["sk_rev_mins"][a:b] == sk["rev_mins"][sk["rev_offs"][i]:sk["rev_offs"][i + 1]]).all()
        a, b = golden["sk_hpc_offs"][slot], golden["sk_hpc_offs"][slot + 1]
        assert (golden["sk_hpc_seq"][a:b] == sk["hpc_seq"][sk["hpc_offs"][i]:sk["hpc_offs"][i + 1]]).all()
        assert (golden["sk_hpc_qual"][a:b] == sk["hpc_qual"][sk["hpc_offs"][i]:sk["hpc_offs"][i + 1]]).all()
        assert golden["sk_hpc_err"][slot] == sk["hpc_err"][i]


def test_golden_alignments(port, golden):
    seq, offs, k = golden["seq"], golden["offs"], int(golden["k"])
    for t in range(len(golden["aln_i"])):
        i, j = int(golden["aln_i"][t]), int(golden["aln_j"][t])
        a, b = seq[offs[i]:offs[i + 1]], seq[offs[j]:offs[j + 1]]
        exp = golden["aln_res"][t]
        if exp[5]:
            b = port.revcomp(b)
        r = port.align(a, b, int(golden["aln_go"][t]), float(golden["aln_e"][t]), k)
        assert [r["aligned"], r["score"], r["end_query"], r["end_ref"], r["comp_len"]] == exp[:5].tolist(), t


def _assign_from_ref(sizes, ids, strands):
    ra, pos = {}, 0
    for c, sz in enumerate(sizes):
        for j in range(1, sz):
            ra[int(ids[pos + j])] = (c, int(strands[pos + j]))
        pos += sz
    return ra


@pytest.mark.parametrize("mode,name", [(0, "sahlin"), (1, "fast"), (2, "furious")])
def test_golden_clusters(port, golden, mode, name):
    seq, qual, offs = golden["seq"], golden["qual"], golden["offs"]
    k, w = int(golden["k"]), int(golden["w"])
    score, err = port.fill_scores(qual, offs, k)
    order = port.sort_order(score)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    status = sk["status"].copy()
    for i in range(len(err)):
        if port.L.orc_quality_filtered(err[i], 7.0):
            status[i] = 0
    pc = port.cluster(seq, offs, sk, score, err, order, k, w, mode, pemp.load_p_shared(k, w), status=status)
    ra = _assign_from_ref(golden[f"cl_{name}_sizes"], golden[f"cl_{name}_ids"], golden[f"cl_{name}_strands"])
    pa = {int(order[s]): (int(pc["cluster"][s]), int(pc["strand"][s])) for s in range(len(order)) if pc["cluster"][s] >= 0}
    assert pa == ra
    assert pc["n_aligned"] == int(golden[f"cl_{name}_n_aligned"])


# ---------------------------------------------------------------- (3) live differential against the reference
def test_ref_differential_sketch(port, ref):
    seq, qual, offs = synth.make_reads(150, seed=21)
    rng = np.random.default_rng(1)
    for i in range(150):
        s, q = seq[offs[i]:offs[i + 1]], qual[offs[i]:offs[i + 1]]
        assert port.qual_score(q, 11) == ref.qual_score(q, 11)
        assert port.error_rate(q) == ref.error_rate(q)
        hs, hq = port.hpc(s, q)
        rs, rq = ref.hpc(s, q)
        assert (hs == rs).all() and (hq == rq).all()
        k = int(rng.integers(2, 32))
        w = k + int(rng.integers(0, 12))
        kv = port.kmer_encode(hs, k)
        assert (kv == ref.kmer_encode(hs, k)).all()
        if len(kv) >= w - k + 1:
            assert (port.minimizers(kv, k, w) == ref.minimizers(kv, k, w)).all()
        assert (port.revcomp(hs) == ref.revcomp(hs)).all()


def test_ref_differential_align(port, ref):
    seq, _, offs = synth.make_reads(60, seed=22, reads_per_gene=6, min_len=200, max_len=500)
    rng = np.random.default_rng(2)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for t in range(60):
        i, j = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        a, b = seq[offs[i]:offs[i + 1]].copy(), seq[offs[j]:offs[j + 1]].copy()
        if t % 4 == 0:  # homopolymer tie storms
            a = np.repeat(acgt[rng.integers(0, 4, 40)], rng.integers(1, 9, 40))
            b = np.repeat(acgt[rng.integers(0, 4, 40)], rng.integers(1, 9, 40))
        if t % 4 == 1:  # containment
            b = a[int(rng.integers(0, 50)):len(a) - int(rng.integers(0, 50))].copy()
        go, e = int(rng.integers(2, 6)), float(rng.choice([0.02, 0.1, 0.2, 0.3]))
        x, y = port.align(a, b, go, e, 11, want_comp=True), ref.align(a, b, go, e, 11, want_comp=True)
        for key in ("aligned", "score", "end_query", "end_ref", "comp_len", "comp"):
            assert x[key] == y[key], (t, key)


def test_ref_differential_mapped_ratio(port, ref):
    k, w = 11, 15
    tab = pemp.load_p_shared(k, w)
    nmax = port.nmax_table(tab, 0.1)
    rng = np.random.default_rng(3)
    for _ in range(300):
        n_mins = int(rng.integers(5, 400))
        nh = int(rng.integers(1, n_mins + 1))
        idx = np.sort(rng.choice(n_mins, nh, replace=False)).astype(np.uint32)
        pos = (idx * 3 + rng.integers(0, 3, nh)).astype(np.uint32)
        hpc_len = int(pos[-1]) + int(rng.integers(1, 40))
        e1, e2 = float(rng.uniform(0.001, 0.2)), float(rng.uniform(0.001, 0.2))
        c1, c2 = port.L.orc_err_class(e1), port.L.orc_err_class(e2)
        assert tab[c1, c2] == ref.pmin_shared(k, w, e1, e2)
        a = port.mapped_ratio(hpc_len, n_mins, pos, idx, tab[c1, c2], 0.1)
        b = ref.mapped_ratio(k, w, hpc_len, e2, e1, n_mins, pos, idx, 0.1)
        assert a == b
        assert port.total_mapped(hpc_len, n_mins, pos, idx, nmax[c1, c2]) / hpc_len == b


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_ref_differential_cluster(port, ref, mode):
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(160, seed=23 + mode, reads_per_gene=10, min_len=300, max_len=700)
    rb = ref.sort_and_sketch(seq, qual, offs, k, w)
    score, err, order = rb.scores()
    rc = rb.cluster(mode)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    status = sk["status"].copy()
    for i in range(len(err)):
        if port.L.orc_quality_filtered(err[i], 7.0):
            status[i] = 0
    pc = port.cluster(seq, offs, sk, score, err, order, k, w, mode, pemp.load_p_shared(k, w), status=status)
    ra = _assign_from_ref(rc["sizes"], rc["ids"], rc["strands"])
    pa = {int(order[s]): (int(pc["cluster"][s]), int(pc["strand"][s])) for s in range(len(order)) if pc["cluster"][s] >= 0}
    assert pa == ra and pc["n_aligned"] == rc["n_aligned"] and pc["n_clusters"] == rc["n_clusters"]


# ---------------------------------------------------------------- a12: UpdateMinDB (minimizer.cpp:124-160)
def _update_case(port):
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(60, seed=77, reads_per_gene=6, min_len=300, max_len=700)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    fwd = lambda r: sk["mins"][sk["min_offs"][r]:sk["min_offs"][r + 1]]
    reps = [0, 6, 12, 18, 24, 30, 36, 1, 42]            # clusters 0..8 (cluster 7 shares a gene with cluster 0)
    upd = [(2, 13), (5, 50), (2, 14), (0, 1), (8, 42), (3, 55), (3, 18)]  # same gene, foreign gene, twice, no-op, there and back
    cm = np.concatenate([fwd(r) for r in reps])
    co = np.cumsum([0] + [len(fwd(r)) for r in reps]).astype(np.int64)
    um = np.concatenate([fwd(r) for _, r in upd])
    uo = np.cumsum([0] + [len(fwd(r)) for _, r in upd]).astype(np.int64)
    final = list(reps)
    for c, r in upd:
        final[c] = r
    fm = np.concatenate([fwd(r) for r in final])
    fo = np.cumsum([0] + [len(fwd(r)) for r in final]).astype(np.int64)
    return (cm, co, [c for c, _ in upd], um, uo), (fm, fo)


def _drop_empty(db):
    keys, offs, cls = db
    return {int(k): cls[offs[i]:offs[i + 1]].tolist() for i, k in enumerate(keys) if offs[i + 1] > offs[i]}


def test_update_mindb_port_equals_fresh_build(port):
    """After UpdateMinDB the non-empty posting lists are those of a MinDB built from the final representatives —
    the property the device index update (icb_clusterer_update_representative) is tested through."""
    case, (fm, fo) = _update_case(port)
    got = port.mindb_after_updates(*case)
    fresh = port.mindb_after_updates(fm, fo, [], np.zeros(0, MIN_DT), np.zeros(1, np.int64))
    assert _drop_empty(got) == _drop_empty(fresh)
    assert len(got[0]) > len(fresh[0])  # emptied keys stay (minimizer.cpp:150-152)


def test_update_mindb_port_vs_reference(port, ref):
    case, (fm, fo) = _update_case(port)
    r, p = ref.mindb_after_updates(*case), port.mindb_after_updates(*case)
    for a, b in zip(r, p):
        assert (a == b).all()
    fresh = ref.mindb_after_updates(fm, fo, [], np.zeros(0, MIN_DT), np.zeros(1, np.int64))
    assert _drop_empty(r) == _drop_empty(fresh)


def test_update_mindb_golden(port):
    """The port against the fixture written by the reference's own UpdateMinDB (tests/golden/make_golden_update_mindb.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_update_mindb.npz"))
    keys, offs, cls = port.mindb_after_updates(g["cls_mins"], g["cls_offs"], g["upd_cls"], g["upd_mins"], g["upd_offs"])
    assert (keys == g["keys"]).all() and (offs == g["offs"]).all() and (cls == g["cls"]).all()
    assert (offs[1:] == offs[:-1]).any()  # emptied keys are kept


def test_cluster_seeded_port_properties(port):
    """orc_cluster_seeded (the oracle of the device a12 test): without seeds it is orc_cluster_reads; seeding the final
    representatives directly equals seeding the old ones and applying the UpdateMinDB restatement."""
    from isonclust2_b200 import pemp, synth
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(160, seed=77, reads_per_gene=8, min_len=300, max_len=700)
    score, err = port.fill_scores(qual, offs, k)
    order = port.sort_order(score)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    ps = pemp.load_p_shared(k, w)
    for mode in (0, 1):
        a = port.cluster(seq, offs, sk, score, err, order, k, w, mode, ps)
        b = port.cluster_seeded(seq, offs, sk, score, err, [], [], order, k, w, mode, ps)
        assert a["n_clusters"] == b["n_clusters"] and (a["cluster"] == b["cluster"]).all() and (a["strand"] == b["strand"]).all()
        assert (order[a["founder"]] == b["founder"]).all() and a["n_aligned"] == b["n_aligned"]
        good = [int(r) for r in order if sk["status"][r] == 1]
        reps = good[0:20:2]
        upd = [(2, good[21]), (5, good[23]), (2, good[25]), (0, reps[0])]
        final = list(reps)
        for c, r in upd:
            final[c] = r
        used = set(reps) | {r for _, r in upd}
        q = [r for r in good if r not in used]
        u = port.cluster_seeded(seq, offs, sk, score, err, reps, upd, q, k, w, mode, ps)
        f = port.cluster_seeded(seq, offs, sk, score, err, final, [], q, k, w, mode, ps)
        assert u["n_clusters"] == f["n_clusters"] and (u["cluster"] == f["cluster"]).all() and (u["strand"] == f["strand"]).all()
        assert (u["founder"] == f["founder"]).all() and u["n_clusters"] > len(reps)
