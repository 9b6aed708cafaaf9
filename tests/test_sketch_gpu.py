"""GPU parity: stage 1 (quality scores + sketches) through the C ABI vs the oracle, bit-exact."""
import numpy as np
import pytest

from isonclust2_b200 import api, synth

pytestmark = pytest.mark.gpu


def _flat(reads, quals=None):
    seq = np.frombuffer(b"".join(reads), np.uint8).copy()
    if quals is None:
        qual = np.full(len(seq), ord("5"), np.uint8)
    else:
        qual = np.frombuffer(b"".join(quals), np.uint8).copy()
    offs = np.cumsum([0] + [len(r) for r in reads]).astype(np.int64)
    return seq, qual, offs


def _check_sketch(port, ctx, seq, qual, offs, k, w, skip=None):
    rs = ctx.upload(seq, qual, offs)
    try:
        sk = rs.sketch(k, w, skip=skip)
    finally:
        rs.close()
    ref = port.sketch_batch(seq, qual, offs, k, w)
    n = len(offs) - 1
    for i in range(n):
        if skip is not None and skip[i]:
            assert sk["status"][i] == 0 and sk["min_offs"][i + 1] == sk["min_offs"][i]
            continue
        assert sk["status"][i] == ref["status"][i], (i, sk["status"][i], ref["status"][i])
        if ref["status"][i] != 1:
            assert sk["min_offs"][i + 1] == sk["min_offs"][i] and sk["rev_offs"][i + 1] == sk["rev_offs"][i]
            continue
        lh = ref["hpc_offs"][i + 1] - ref["hpc_offs"][i]
        assert sk["hpc_len"][i] == lh, i
        a = int(offs[i])
        assert (sk["hpc_seq"][a:a + lh] == ref["hpc_seq"][ref["hpc_offs"][i]:ref["hpc_offs"][i + 1]]).all(), i
        assert (sk["hpc_qual"][a:a + lh] == ref["hpc_qual"][ref["hpc_offs"][i]:ref["hpc_offs"][i + 1]]).all(), i
        assert sk["hpc_err"][i] == ref["hpc_err"][i], i
        m = sk["mins"][sk["min_offs"][i]:sk["min_offs"][i + 1]]
        e = ref["mins"][ref["min_offs"][i]:ref["min_offs"][i + 1]]
        assert len(m) == len(e) and (m == e).all(), (i, "fwd")
        m = sk["rev_mins"][sk["rev_offs"][i]:sk["rev_offs"][i + 1]]
        e = ref["rev_mins"][ref["rev_offs"][i]:ref["rev_offs"][i + 1]]
        assert len(m) == len(e) and (m == e).all(), (i, "rev")
    return sk


def test_scores_match_oracle(port, ctx):
    seq, qual, offs = synth.make_reads(3000, seed=31)
    rs = ctx.upload(seq, qual, offs)
    for k in (11, 13, 16, 17, 31, 3):
        score, err = rs.score(k)
        ps, pe = port.fill_scores(qual, offs, k)
        assert (score == ps).all(), k
        assert (err == pe).all(), k
    rs.close()
    assert (api.init_qual_tab() == port.tab).all() and (api.init_qual_tab_nomin() == port.tab_nomin).all()


def test_scores_golden_and_order(port, ctx, golden):
    seq, qual, offs, k = golden["seq"], golden["qual"], golden["offs"], int(golden["k"])
    rs = ctx.upload(seq, qual, offs)
    score, err = rs.score(k)
    rs.close()
    assert (score == golden["score"]).all() and (err == golden["err"]).all()
    assert (api.sort_by_qual_scores(score) == golden["order"]).all()


def test_scores_edge_lengths(port, ctx):
    # reads of length 0, 1, k, 2k, 2k+1, and unaligned starts; wide Phred range (ratio-table fallback above Q63)
    rng = np.random.default_rng(4)
    k = 11
    lens = [0, 1, 5, 11, 22, 23, 24, 37, 38, 39, 40, 41, 100, 257, 1000, 3]
    reads = [bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8)) for n in lens]
    quals = [bytes(rng.integers(33, 127, n).astype(np.uint8)) for n in lens]
    seq, qual, offs = _flat(reads, quals)
    rs = ctx.upload(seq, qual, offs)
    score, err = rs.score(k)
    rs.close()
    ps, pe = port.fill_scores(qual, offs, k)
    assert (score == ps).all() and (err == pe).all()


@pytest.mark.parametrize("k,w", [(11, 15), (13, 20), (10, 10), (16, 25), (17, 21), (21, 30), (31, 40), (2, 4), (3, 5), (11, 60)])
def test_sketch_matches_oracle(port, ctx, k, w):
    seq, qual, offs = synth.make_reads(400, seed=32 + k)
    _check_sketch(port, ctx, seq, qual, offs, k, w)


def test_sketch_golden(port, ctx, golden):
    seq, qual, offs = golden["seq"], golden["qual"], golden["offs"]
    k, w = int(golden["k"]), int(golden["w"])
    rs = ctx.upload(seq, qual, offs)
    score, err = rs.score(k)
    skip = api.quality_filtered(err, 7.0)
    sk = rs.sketch(k, w, skip=skip)
    rs.close()
    order = api.sort_by_qual_scores(score)
    for slot, i in enumerate(order):
        gs = int(golden["sk_status"][slot])
        st = int(sk["status"][i])
        assert (st, gs) in ((0, 0), (1, 1), (2, 2), (3, 0)), (i, st, gs)
        if st != 1:
            continue
        a, b = golden["sk_min_offs"][slot], golden["sk_min_offs"][slot + 1]
        assert (golden["sk_mins"][a:b] == sk["mins"][sk["min_offs"][i]:sk["min_offs"][i + 1]]).all()
        a, b = golden["sk_rev_offs"][slot], golden["sk_rev_offs"][slot + 1]
        assert (golden["sk_rev_mins"][a:b] == sk["rev_mins"][sk["rev_offs"][i]:sk["rev_offs"][i + 1]]).all()
        a, b = golden["sk_hpc_offs"][slot], golden["sk_hpc_offs"][slot + 1]
        o = int(offs[i])
        assert (golden["sk_hpc_seq"][a:b] == sk["hpc_seq"][o:o + (b - a)]).all()
        assert (golden["sk_hpc_qual"][a:b] == sk["hpc_qual"][o:o + (b - a)]).all()
        assert golden["sk_hpc_err"][slot] == sk["hpc_err"][i]


def test_sketch_edge_cases(port, ctx):
    rng = np.random.default_rng(6)
    k, w = 11, 15
    rnd = lambda n: bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
    reads = [
        b"", b"A", rnd(14), rnd(15), rnd(22), rnd(23),	    # too short / boundary lengths
        b"A" * 500,					    # one run -> compressed length 1
        b"AC" * 11, b"AC" * 12,				    # compressed length == 2k and just above
        b"AAAACCCCGGGGTTTT" * 40,			    # runs of 4
        rnd(2047), rnd(2048), rnd(2049), rnd(4095), rnd(4096),  # tile boundaries of the shared-memory path
        b"A" * 2047 + rnd(300),				    # a run that spans a whole staging tile
        rnd(1000) + b"G" * 3000 + rnd(90),		    # long run crossing tiles, still on the fast path
        rnd(4097), rnd(9000), rnd(20000),		    # global-scratch path
        b"T" * 5000 + rnd(100),
        rnd(3071), rnd(3072), rnd(3073),		    # limit of the warp-per-read path
        rnd(7) + b"C" * 9 + rnd(16) + b"G" * 33 + rnd(5) + b"T" * 16 + rnd(11) + b"A" * 600 + rnd(700),  # runs across 16-base chunks
        b"AC" * 1536, b"ACG" * 1000,			    # as many minimizers as a read can have
    ]
    quals = [bytes(rng.integers(35, 75, len(r)).astype(np.uint8)) for r in reads]
    seq, qual, offs = _flat(reads, quals)
    _check_sketch(port, ctx, seq, qual, offs, k, w)
    skip = np.zeros(len(reads), np.uint8)
    skip[[3, 9, 18]] = 1
    _check_sketch(port, ctx, seq, qual, offs, k, w, skip=skip)
    _check_sketch(port, ctx, seq, qual, offs, 13, 20)
    # unaligned slot starts: the same reads behind a 1..15-byte first read
    for pad in (1, 5, 15):
        s2, q2, o2 = _flat([rnd(pad)] + reads[6:], [bytes([60] * pad)] + quals[6:])
        _check_sketch(port, ctx, s2, q2, o2, k, w)


def test_sketch_empty_and_invalid(ctx):
    rs = ctx.upload(np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.int64))
    sk = rs.sketch(11, 15)
    assert len(sk["mins"]) == 0 and len(sk["status"]) == 0
    rs.close()
    # a non-ACGT base in a read that reaches RevComp -> the reference throws (util.cpp:31-33)
    seq, qual, offs = _flat([b"ACGTTGCA" * 10, b"ACGTNGCA" * 10])
    rs = ctx.upload(seq, qual, offs)
    with pytest.raises(api.IcbError) as ei:
        rs.sketch(11, 15)
    assert ei.value.code == -3
    # ... but not if the host's quality filter already skipped it
    sk = rs.sketch(11, 15, skip=np.array([0, 1], np.uint8))
    assert list(sk["status"]) == [1, 0]
    rs.close()
    with pytest.raises(api.IcbError):
        ctx.upload(seq, qual, offs).sketch(11, 10)  # w < k


def test_sketch_full_size_properties(ctx):
    """BASELINE-size property check (no oracle): 200k reads, invariants of the minimizer stream."""
    k, w = 11, 15
    seq, qual, offs = synth.make_reads(200_000, seed=33, err_set=(0.08,))
    rs = ctx.upload(seq, qual, offs)
    sk = rs.sketch(k, w, hpc=True)
    rs.close()
    W = w - k + 1
    assert (sk["status"] == 1).all()
    lens = offs[1:] - offs[:-1]
    assert (sk["hpc_len"] <= lens).all() and (sk["hpc_len"] > 0).all()
    for arr, o in ((sk["mins"], sk["min_offs"]), (sk["rev_mins"], sk["rev_offs"])):
        cnt = o[1:] - o[:-1]
        nwin = sk["hpc_len"] - k - W + 1
        assert (cnt >= 1).all() and (cnt <= nwin).all()
        first = o[:-1]
        assert (arr["Index"][first] == 0).all()
        inner = np.ones(len(arr), bool)
        inner[first] = False
        assert (arr["Index"][inner] == arr["Index"][np.flatnonzero(inner) - 1] + 1).all()  # ranks count up
        assert (arr["Pos"][inner] > arr["Pos"][np.flatnonzero(inner) - 1]).all()	    # positions increase
        assert (np.diff(arr["Pos"])[inner[1:]] <= W).all()				    # consecutive minimizers within one window
        assert (arr["Min"] < 4 ** k).all()
    # compressed sequence has no equal neighbours inside a read
    hs = sk["hpc_seq"]
    for i in range(0, 200_000, 997):
        a = int(offs[i]); h = hs[a:a + sk["hpc_len"][i]]
        assert (h[1:] != h[:-1]).all()
