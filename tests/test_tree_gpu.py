"""GPU: the in-process merge tree on resident clusterers (icb_cluster_tree, configs[4]) against the UNMODIFIED reference
running the same tree as `cluster -l A -r B` merges (cluster.cpp:67-322)."""
import numpy as np
import pytest

from isonclust2_b200 import api, synth

pytestmark = pytest.mark.gpu


def _members(rc):
    ra, pos = {}, 0
    for c, sz in enumerate(rc["sizes"]):
        for j in range(1, sz):
            ra[int(rc["ids"][pos + j])] = (c, int(rc["strands"][pos + j]))
        pos += sz
    return ra


@pytest.mark.parametrize("mode,n_batches", [(0, 5), (1, 4), (0, 1)])
def test_tree_matches_reference_tree(ref, ctx, mode, n_batches):
    k, w = 11, 15
    per = 420
    n = per * n_batches
    seq, qual, offs = synth.make_reads(n, seed=401 + mode, reads_per_gene=30)
    rb = ref.sort_and_sketch(seq, qual, offs, k, w)
    _, _, order = rb.scores()
    nodes = [rb.slice(b * per, (b + 1) * per - 1, b) for b in range(n_batches)]
    last = None
    for nd in nodes:
        last = nd.cluster(mode)
    level = list(range(n_batches))
    aligned = 0
    while len(level) > 1:  # adjacent pairs, an odd last node is carried up
        nxt = []
        for i in range(0, len(level) - 1, 2):
            last = nodes[level[i]].cluster(mode, right=nodes[level[i + 1]])
            nxt.append(level[i])
        if len(level) % 2:
            nxt.append(level[-1])
        level = nxt
    want = _members(last)

    rs = ctx.upload(seq, qual, offs)
    try:
        score, err = rs.score(k)
        assert (api.sort_by_qual_scores(score) == order).all()
        rs.sketch_async(k, w, skip=api.quality_filtered(err, 7.0))
        bounds = np.arange(0, n + 1, per, dtype=np.int64)
        out = {}
        for nw in (1, 3):
            # the reference's `sort` default MinClsSize = 3 (args.h:21) is what its batches carry: small clusters of merged
            # right nodes are dropped
            cl, st, ncl, stats = api.cluster_tree(rs, order, bounds, k, w, n_workers=nw, min_cls_size=3, mode=mode, chunk_size=96)
            out[nw] = (cl.copy(), st.copy(), ncl)
            assert stats["tree_merges"] == n_batches - 1
        assert (out[1][0] == out[3][0]).all() and (out[1][1] == out[3][1]).all() and out[1][2] == out[3][2]
        cl, st, ncl = out[3]
        assert ncl == last["n_clusters"]
        got = {int(order[s]): (int(cl[s]), int(st[s])) for s in range(n) if cl[s] >= 0}
        assert got == want
    finally:
        rs.close()
