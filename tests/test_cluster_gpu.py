"""GPU parity: stage 2 (device index + candidate query) and the host greedy loop vs the oracle — identical cluster
assignment, strands, founders and alignment-invocation counts, in all three modes and for every chunk size."""
import numpy as np
import pytest

from isonclust2_b200 import api, pemp, synth

pytestmark = pytest.mark.gpu


def _gpu_cluster(ctx, seq, qual, offs, k, w, mode, chunk_size, **kw):
    rs = ctx.upload(seq, qual, offs)
    try:
        score, err = rs.score(k)
        skip = api.quality_filtered(err, kw.get("min_qual", 7.0))
        rs.sketch_async(k, w, skip=skip)
        order = api.sort_by_qual_scores(score)
        cl = api.Clusterer(rs, k, w, mode=mode, chunk_size=chunk_size, **kw)
        try:
            c, s = cl.cluster(order)
            return dict(order=order, cluster=c, strand=s, founders=cl.founders(), stats=cl.stats(), n_clusters=cl.n_clusters)
        finally:
            cl.close()
    finally:
        rs.close()


def _port_cluster(port, seq, qual, offs, k, w, mode, **kw):
    score, err = port.fill_scores(qual, offs, k)
    order = port.sort_order(score)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    status = sk["status"].copy()
    for i in range(len(err)):
        if port.L.orc_quality_filtered(err[i], kw.get("min_qual", 7.0)):
            status[i] = 0
    pc = port.cluster(seq, offs, sk, score, err, order, k, w, mode, pemp.load_p_shared(k, w), status=status, **kw)
    pc["order"] = order
    return pc


def _same(g, p):
    # decided on the device (the host loop only takes chunks the device loop hands back: none in these cases)
    assert g["stats"]["dl_chunks"] == g["stats"]["chunks"] > 0 and g["stats"]["dl_fallback_chunks"] == 0, g["stats"]
    assert (g["order"] == p["order"]).all()
    assert g["n_clusters"] == p["n_clusters"]
    assert (g["cluster"] == p["cluster"]).all(), np.flatnonzero(g["cluster"] != p["cluster"])[:10]
    assert (g["strand"] == p["strand"]).all()
    # founders: the oracle reports sorted-slot indices, the product read ids
    assert (g["founders"] == g["order"][p["founder"]]).all()
    assert g["stats"]["aln_invoked"] == p["n_aligned"]


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("chunk_size", [1, 7, 64, 256])
def test_cluster_matches_oracle(port, ctx, mode, chunk_size):
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(220, seed=51, reads_per_gene=11, min_len=300, max_len=800)
    p = _port_cluster(port, seq, qual, offs, k, w, mode)
    g = _gpu_cluster(ctx, seq, qual, offs, k, w, mode, chunk_size)
    _same(g, p)


@pytest.mark.parametrize("mode,name", [(0, "sahlin"), (1, "fast"), (2, "furious")])
def test_cluster_golden(ctx, golden, mode, name):
    """Fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
    seq, qual, offs = golden["seq"], golden["qual"], golden["offs"]
    k, w = int(golden["k"]), int(golden["w"])
    g = _gpu_cluster(ctx, seq, qual, offs, k, w, mode, 32)
    sizes, ids, strands = golden[f"cl_{name}_sizes"], golden[f"cl_{name}_ids"], golden[f"cl_{name}_strands"]
    ra, pos = {}, 0
    for c, sz in enumerate(sizes):
        for j in range(1, sz):
            ra[int(ids[pos + j])] = (c, int(strands[pos + j]))
        pos += sz
    ga = {int(g["order"][s]): (int(g["cluster"][s]), int(g["strand"][s])) for s in range(len(g["order"])) if g["cluster"][s] >= 0}
    assert ga == ra
    assert g["stats"]["aln_invoked"] == int(golden[f"cl_{name}_n_aligned"])


@pytest.mark.parametrize("mode", [0, 1])
def test_cluster_vs_reference_batch(ref, ctx, mode):
    """A 1500-read batch against the UNMODIFIED reference's ClusterSortedReads (ties, many clusters in fast mode)."""
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(1500, seed=52 + mode)
    rb = ref.sort_and_sketch(seq, qual, offs, k, w)
    _, _, order = rb.scores()
    rc = rb.cluster(mode)
    g = _gpu_cluster(ctx, seq, qual, offs, k, w, mode, 256)
    assert (g["order"] == order).all()
    ra, pos = {}, 0
    for c, sz in enumerate(rc["sizes"]):
        for j in range(1, sz):
            ra[int(rc["ids"][pos + j])] = (c, int(rc["strands"][pos + j]))
        pos += sz
    ga = {int(g["order"][s]): (int(g["cluster"][s]), int(g["strand"][s])) for s in range(len(order)) if g["cluster"][s] >= 0}
    assert ga == ra
    assert g["n_clusters"] == rc["n_clusters"] and g["stats"]["aln_invoked"] == rc["n_aligned"]


def test_cluster_params_variants(port, ctx):
    k, w = 13, 20
    seq, qual, offs = synth.make_reads(150, seed=54, reads_per_gene=10, min_len=300, max_len=700)
    for kw in (dict(min_shared=3, min_frac=0.5), dict(mapped_th=0.9, aligned_th=0.4), dict(min_prob_no_hits=0.3, min_qual=12.0)):
        p = _port_cluster(port, seq, qual, offs, k, w, 0, **kw)
        gk = {{"mapped_th": "mapped_threshold", "aligned_th": "aligned_threshold", "min_frac": "min_fraction"}.get(a, a): b
              for a, b in kw.items()}
        g = _gpu_cluster(ctx, seq, qual, offs, k, w, 0, 48, **gk)
        _same(g, p)


def test_cluster_no_pemp_row(ctx):
    seq, qual, offs = synth.make_reads(20, seed=55, min_len=300, max_len=400)
    rs = ctx.upload(seq, qual, offs)
    rs.score(31)
    rs.sketch_async(31, 40)  # the empirical table covers k = 10..30 only
    with pytest.raises(api.IcbError) as ei:
        api.Clusterer(rs, 31, 40, mode="sahlin", p_shared=pemp.load_p_shared(31, 40))
    assert ei.value.code == -4
    rs.close()


def _members(rc):
    """read -> (cluster, strand) from the reference's cluster lists (element 0 of a cluster is the representative copy)."""
    ra, pos = {}, 0
    for c, sz in enumerate(rc["sizes"]):
        for j in range(1, sz):
            ra[int(rc["ids"][pos + j])] = (c, int(rc["strands"][pos + j]))
        pos += sz
    return ra


@pytest.mark.parametrize("mode", [0, 1])
def test_cluster_batches_and_merge_vs_reference(ref, ctx, mode):
    """Two sorted batches clustered concurrently (icb_cluster_batches), then merged left <- right (icb_clusterer_seed +
    icb_cluster_reads over the right batch's representatives) against the UNMODIFIED reference doing the same
    (`cluster -l A`, `cluster -l B`, `cluster -l A -r B`, cluster.cpp:67-322)."""
    k, w = 11, 15
    n, h = 1400, 800
    seq, qual, offs = synth.make_reads(n, seed=61 + mode)
    rb = ref.sort_and_sketch(seq, qual, offs, k, w)
    _, _, order = rb.scores()
    A, B = rb.slice(0, h - 1, 0), rb.slice(h, n - 1, 1)
    ra, rbb = A.cluster(mode), B.cluster(mode)
    mem_a, mem_b = _members(ra), _members(rbb)
    rm = A.cluster(mode, right=B)
    mem_m = _members(rm)

    rs = ctx.upload(seq, qual, offs)
    try:
        score, err = rs.score(k)
        assert (api.sort_by_qual_scores(score) == order).all()
        rs.sketch_async(k, w, skip=api.quality_filtered(err, 7.0))
        bounds = np.array([0, h, n], np.int64)
        res = {}
        for nw in (1, 2):
            cl, st, ncl, stats = api.cluster_batches(rs, order, bounds, k, w, n_workers=nw, mode=mode, chunk_size=128)
            res[nw] = (cl.copy(), st.copy(), ncl.copy())
        assert (res[1][0] == res[2][0]).all() and (res[1][1] == res[2][1]).all() and (res[1][2] == res[2][2]).all()
        cl, st, ncl = res[2]
        assert int(ncl[0]) == ra["n_clusters"] and int(ncl[1]) == rbb["n_clusters"]
        ga = {int(order[s]): (int(cl[s]), int(st[s])) for s in range(0, h) if cl[s] >= 0}
        gb = {int(order[s]): (int(cl[s]), int(st[s])) for s in range(h, n) if cl[s] >= 0}
        assert ga == mem_a and gb == mem_b
        # founders in cluster-id order = first read assigned to each cluster
        def founders(lo, hi, nc):
            f = np.full(nc, -1, np.int32)
            for s in range(lo, hi):
                if cl[s] >= 0 and f[cl[s]] < 0:
                    f[cl[s]] = order[s]
            return f
        fa, fb = founders(0, h, int(ncl[0])), founders(h, n, int(ncl[1]))
        m = api.Clusterer(rs, k, w, mode=mode, chunk_size=64)
        try:
            m.seed(fa)
            tc, ts = m.cluster(fb)
            assert m.n_clusters == rm["n_clusters"]
        finally:
            m.close()
        gm = dict(ga)
        for r, (c, s) in gb.items():
            assert tc[c] >= 0
            gm[r] = (int(tc[c]), s * (-1 if ts[c] == -1 else 1))
        assert gm == mem_m
    finally:
        rs.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_update_representative_equals_fresh_seed(ctx, mode):
    """a12, UpdateMinDB (minimizer.cpp:124-160) on the device index: seeding clusters and then swapping representatives gives
    the assignments of a clusterer seeded with the final representatives directly (the MinDB equality the CPU suite pins on
    the reference, tests/test_oracle.py::test_update_mindb_port_vs_reference), including for clusters founded later."""
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(400, seed=91 + mode, reads_per_gene=8, min_len=300, max_len=900)
    rs = ctx.upload(seq, qual, offs)
    try:
        score, err = rs.score(k)
        rs.sketch_async(k, w, skip=api.quality_filtered(err, 7.0))
        order = api.sort_by_qual_scores(score)
        ok = rs.sketch_download(hpc=False)["status"] == 1
        good = [int(r) for r in order if ok[r]]
        reps = good[0:40:2]                         # 20 clusters
        upd = [(3, good[41]), (7, good[43]), (3, good[45]), (0, reps[0]), (11, good[47]), (11, reps[11]), (19, good[49])]
        final = list(reps)
        for c, r in upd:
            final[c] = r
        used = set(reps) | {r for _, r in upd}
        queries = np.array([r for r in good if r not in used], np.int32)
        out = {}
        for name in ("updated", "fresh"):
            cl = api.Clusterer(rs, k, w, mode=mode, chunk_size=32)
            try:
                if name == "updated":
                    cl.seed(reps)
                    for c, r in upd:
                        cl.update_representative(c, r)
                    assert cl.founders().tolist() == final
                else:
                    cl.seed(final)
                c1, s1 = cl.cluster(queries)
                out[name] = (c1, s1, cl.n_clusters, cl.founders())
            finally:
                cl.close()
        a, b = out["updated"], out["fresh"]
        assert a[2] == b[2] and (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[3] == b[3]).all()
        assert (a[0] < len(reps)).any() and a[2] > len(reps)  # reads joined seeded clusters and founded new ones
        # a swapped-in representative attracts its own gene: some query lands in an updated cluster
        assert np.isin(a[0], [3, 7, 19]).any()
        with pytest.raises(api.IcbError):
            m = api.Clusterer(rs, k, w, mode=mode)
            try:
                m.seed(reps[:2])
                m.update_representative(5, reps[0])
            finally:
                m.close()
    finally:
        rs.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_update_representative_vs_oracle(port, ctx, mode):
    """a12 against the ORACLE: clusters seeded, representatives swapped with icb_clusterer_update_representative (the device
    UpdateMinDB, minimizer.cpp:124-160 + cluster.cpp:279-291), then reads clustered — cluster / strand per read, founders and
    the alignment count equal the oracle's loop run on a MinDB updated by its UpdateMinDB restatement (which
    tests/test_oracle.py pins on the reference's own UpdateMinDB)."""
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(400, seed=191 + mode, reads_per_gene=8, min_len=300, max_len=900)
    score, err = port.fill_scores(qual, offs, k)
    order = port.sort_order(score)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    status = sk["status"].copy()
    for i in range(len(err)):
        if port.L.orc_quality_filtered(err[i], 7.0):
            status[i] = 0
    good = [int(r) for r in order if status[r] == 1]
    reps = good[0:40:2]
    upd = [(3, good[41]), (7, good[43]), (3, good[45]), (0, reps[0]), (11, good[47]), (11, reps[11]), (19, good[49]), (5, good[51])]
    used = set(reps) | {r for _, r in upd}
    queries = np.array([r for r in good if r not in used], np.int32)
    pc = port.cluster_seeded(seq, offs, sk, score, err, reps, upd, queries, k, w, mode, pemp.load_p_shared(k, w), status=status)
    rs = ctx.upload(seq, qual, offs)
    try:
        s2, e2 = rs.score(k)
        rs.sketch_async(k, w, skip=api.quality_filtered(e2, 7.0))
        cl = api.Clusterer(rs, k, w, mode=mode, chunk_size=32)
        try:
            cl.seed(reps)
            for c, r in upd:
                cl.update_representative(c, r)
            c1, s1 = cl.cluster(queries)
            assert cl.n_clusters == pc["n_clusters"]
            assert (c1 == pc["cluster"]).all(), np.flatnonzero(c1 != pc["cluster"])[:10]
            assert (s1 == pc["strand"]).all()
            assert (cl.founders() == pc["founder"]).all()
            assert cl.stats()["aln_invoked"] == pc["n_aligned"]
            assert np.isin(c1, [3, 7, 19, 5]).any() and cl.n_clusters > len(reps)
        finally:
            cl.close()
    finally:
        rs.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_appended_representatives_vs_oracle(port, ctx, mode):
    """icb_readset_append_from (the consensus path's way of giving a cluster a NEW representative sequence): the same swaps as
    above, but every new representative is a read of a SECOND read set appended to the reserve of the first — the device copies
    its bases and sketches, the caller states its error rates — and the clustering that follows equals the oracle's."""
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(400, seed=191 + mode, reads_per_gene=8, min_len=300, max_len=900)
    score, err = port.fill_scores(qual, offs, k)
    order = port.sort_order(score)
    sk = port.sketch_batch(seq, qual, offs, k, w)
    status = sk["status"].copy()
    for i in range(len(err)):
        if port.L.orc_quality_filtered(err[i], 7.0):
            status[i] = 0
    good = [int(r) for r in order if status[r] == 1]
    reps = good[0:40:2]
    upd = [(3, good[41]), (7, good[43]), (3, good[45]), (11, good[47]), (19, good[49]), (5, good[51])]
    used = set(reps) | {r for _, r in upd}
    queries = np.array([r for r in good if r not in used], np.int32)
    pc = port.cluster_seeded(seq, offs, sk, score, err, reps, upd, queries, k, w, mode, pemp.load_p_shared(k, w), status=status)
    # the second read set: the replacement reads only
    src_ids = [r for _, r in upd]
    s_offs = np.zeros(len(src_ids) + 1, np.int64)
    s_offs[1:] = np.cumsum([offs[r + 1] - offs[r] for r in src_ids])
    s_seq = np.concatenate([seq[offs[r]:offs[r + 1]] for r in src_ids])
    s_qual = np.concatenate([qual[offs[r]:offs[r + 1]] for r in src_ids])
    rs = ctx.upload(seq, qual, offs, reserve=(len(src_ids), int(s_offs[-1]) + 64, 1000))
    src = ctx.upload(s_seq, s_qual, s_offs)
    try:
        s2, e2 = rs.score(k)
        rs.sketch_async(k, w, skip=api.quality_filtered(e2, 7.0))
        src.score(k)
        src.sketch_async(k, w)
        cl = api.Clusterer(rs, k, w, mode=mode, chunk_size=32)
        try:
            cl.seed(reps)
            new_id = {}
            for j, (c, r) in enumerate(upd):
                idx = rs.append_from(src, j, err[r], score[r], sk["hpc_err"][r])
                assert idx == len(offs) - 1 + j
                new_id[idx] = r
                cl.update_representative(c, idx)
            c1, s1 = cl.cluster(queries)
            assert cl.n_clusters == pc["n_clusters"]
            assert (c1 == pc["cluster"]).all(), np.flatnonzero(c1 != pc["cluster"])[:10]
            assert (s1 == pc["strand"]).all()
            f = np.array([new_id.get(int(x), int(x)) for x in cl.founders()])
            assert (f == pc["founder"]).all()
            assert cl.stats()["aln_invoked"] == pc["n_aligned"]
            assert np.isin(c1, [3, 7, 19, 5]).any()
            with pytest.raises(api.IcbError):  # the reserve is used up
                rs.append_from(src, 0, err[upd[0][1]], score[upd[0][1]], sk["hpc_err"][upd[0][1]])
        finally:
            cl.close()
    finally:
        src.close()
        rs.close()


def _per_slot(rc, order):
    """(cluster, strand) per sorted slot from the reference's member lists; -1 / 0 for reads it skipped."""
    slot_of = np.empty(len(order), np.int64)
    slot_of[order] = np.arange(len(order))
    c, s = np.full(len(order), -1, np.int32), np.zeros(len(order), np.int32)
    pos = 0
    for ci, sz in enumerate(rc["sizes"]):
        ids = rc["ids"][pos + 1:pos + sz]
        c[slot_of[ids]] = ci
        s[slot_of[ids]] = rc["strands"][pos + 1:pos + sz]
        pos += sz
    return c, s


@pytest.mark.parametrize("mode,n", [(0, 30000), (1, 15000)])
def test_cluster_benchmark_scale_vs_reference(ref, ctx, mode, n):
    """Parity at the BENCHMARKED scale: one batch of configs[1]'s generator (the first n reads of bench.py's 1 M-read set:
    40 000 genes, 8 % error, < 1 read per gene and batch, so most candidates are random hits tied at small counts and a third
    of the reads go through the reference's candidate-order replay) against the UNMODIFIED reference's ClusterSortedReads.
    sahlin: a full 30 000-read batch; fast: 15 000 reads (the reference needs 5 minutes for 30 000: ~25 000 clusters)."""
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,), n_genes=40000)
    rb = ref.sort_and_sketch(seq, qual, offs, k, w)
    _, _, order = rb.scores()
    rc = rb.cluster(mode)
    g = _gpu_cluster(ctx, seq, qual, offs, k, w, mode, 256)
    assert (g["order"] == order).all()
    rcl, rst = _per_slot(rc, order)
    assert g["n_clusters"] == rc["n_clusters"]
    assert (g["cluster"] == rcl).all(), np.flatnonzero(g["cluster"] != rcl)[:10]
    assert (g["strand"] == rst).all()
    assert g["stats"]["aln_invoked"] == rc["n_aligned"]
    if mode == 0:
        assert g["stats"]["tie_fallbacks"] > n // 20  # the candidate-order path is exercised at this scale
    # decided on the device; the host loop may take the odd chunk in which a read's candidate list exceeds the order kernel's
    # shared-memory plan (reason 2) or the reference's hit map would rehash (reason 3): fast mode against > 10 000 clusters
    assert g["stats"]["dl_chunks"] >= 0.8 * g["stats"]["chunks"] and (g["stats"]["dl_fallback_reasons"] & ~((1 << 2) | (1 << 3))) == 0, g["stats"]


@pytest.mark.parametrize("engine", ["host", "mixed"])
@pytest.mark.parametrize("mode", [0, 1])
def test_cluster_engines_agree_with_reference(ref, ctx, mode, engine, monkeypatch):
    """The host loop (ICB_HOST_LOOP: decisions on the host, the engine the device loop hands unsupported chunks to) and a run in
    which every third chunk is handed back (ICB_DLOOP_FORCE_FALLBACK) against the UNMODIFIED reference on a batch with
    many tied candidates."""
    k, w = 11, 15
    n = 4000
    seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,), n_genes=40000)
    rb = ref.sort_and_sketch(seq, qual, offs, k, w)
    _, _, order = rb.scores()
    rc = rb.cluster(mode)
    if engine == "host":
        monkeypatch.setenv("ICB_HOST_LOOP", "1")
    else:
        monkeypatch.setenv("ICB_DLOOP_FORCE_FALLBACK", "3")
    g = _gpu_cluster(ctx, seq, qual, offs, k, w, mode, 64)
    rcl, rst = _per_slot(rc, order)
    assert g["n_clusters"] == rc["n_clusters"] and (g["cluster"] == rcl).all() and (g["strand"] == rst).all()
    assert g["stats"]["aln_invoked"] == rc["n_aligned"]
    if engine == "host":
        assert g["stats"]["dl_chunks"] == 0
    else:
        assert 0 < g["stats"]["dl_fallback_chunks"] < g["stats"]["chunks"] and g["stats"]["dl_chunks"] > 0


def test_cluster_full_size_properties(ctx):
    """BASELINE-shape property check (no oracle): 120k reads of the configs[1] generator in 4 batches of 30 000, `sahlin`.
    The assignment is a pure function of the sorted order — independent of the chunk size and of how many batches are in
    flight — and carries the greedy loop's invariants."""
    k, w = 11, 15
    n, B = 120_000, 30_000
    seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
    rs = ctx.upload(seq, qual, offs)
    try:
        score, err = rs.score(k)
        filt = api.quality_filtered(err, 7.0)
        rs.sketch_async(k, w, skip=filt)
        order = api.sort_by_qual_scores(score)
        assert (np.diff(score[order]) <= 0).all()  # SortByQualScores: descending, stable
        bounds = np.arange(0, n + 1, B, dtype=np.int64)
        a = api.cluster_batches(rs, order, bounds, k, w, n_workers=4, mode=0, chunk_size=256)
        b = api.cluster_batches(rs, order, bounds, k, w, n_workers=2, mode=0, chunk_size=96)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
        assert a[3]["aln_invoked"] == b[3]["aln_invoked"]
        cl, st, ncl, stats = a
        status = rs.sketch_download(hpc=False)["status"]
        for bi in range(len(bounds) - 1):
            c, s = cl[bounds[bi]:bounds[bi + 1]], st[bounds[bi]:bounds[bi + 1]]
            o = order[bounds[bi]:bounds[bi + 1]]
            live = c >= 0
            assert (live == (status[o] == 1)).all() and ((s == 0) == ~live).all()  # skipped <=> not sketched
            ids = c[live]
            assert ids.max() + 1 == ncl[bi]
            # cluster ids are handed out in order of first appearance, and the first member founds the cluster on strand +1
            first = np.full(int(ncl[bi]), -1, np.int64)
            pos = np.flatnonzero(live)
            firstpos = np.unique(ids, return_index=True)[1]
            first[:] = pos[firstpos]
            assert (np.diff(first) > 0).all()
            assert (s[first] == 1).all()
            assert set(np.unique(s[live]).tolist()) <= {-1, 1}
        # the generator draws ~1 read in 20 per gene family and reverse-complements half: both strands are hit, most reads join
        assert (st == -1).sum() > n // 10 and int(ncl.sum()) < n // 4
        assert stats["aln_invoked"] > 0 and stats["cells_aligned"] > 0
    finally:
        rs.close()
