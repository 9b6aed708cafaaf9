"""GPU parity: stage 3 (semi-global alignment + traceback + k-window count) through the C ABI vs the oracle."""
import numpy as np
import pytest

from isonclust2_b200 import api, synth

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", np.uint8)


def _flat(reads):
    seq = np.frombuffer(b"".join(reads), np.uint8).copy()
    offs = np.cumsum([0] + [len(r) for r in reads]).astype(np.int64)
    return seq, np.full(len(seq), ord("5"), np.uint8), offs


def _check(port, ctx, reads, pairs, k, ref=None):
    seq, qual, offs = _flat(reads)
    rs = ctx.upload(seq, qual, offs)
    try:
        out = rs.align_pairs(pairs, k)
    finally:
        rs.close()
    for t, p in enumerate(pairs):
        a = np.frombuffer(reads[p["read"]], np.uint8)
        b = np.frombuffer(reads[p["rep"]], np.uint8)
        if p["strand"] == -1:
            b = port.revcomp(b)
        # limit -> an e that floors to it: limit = floor((1-e)*k)
        e = 1.0 - (p["limit"] + 0.5) / k
        exp = (ref or port).align(a, b, int(p["gap_open"]), e, k)
        got = out[t]
        assert (got["score"], got["end_query"], got["end_ref"], got["comp_len"]) == \
               (exp["score"], exp["end_query"], exp["end_ref"], exp["comp_len"]), (t, got, exp)
        if exp["comp_len"] >= k:  # shorter than k is UB in the reference (App. C#7)
            assert got["aligned"] == exp["aligned"], (t, got, exp)


def _mk_pairs(rows):
    return np.array(rows, dtype=api.PAIR_DT)


def test_align_synthetic_families(port, ctx):
    seq, _, offs = synth.make_reads(120, seed=41, reads_per_gene=10, min_len=300, max_len=900)
    reads = [seq[offs[i]:offs[i + 1]].tobytes() for i in range(120)]
    rng = np.random.default_rng(7)
    rows = []
    for t in range(150):
        i = int(rng.integers(0, 120))
        j = int(rng.integers(0, 120))
        rows.append((i, j, int(rng.choice([1, -1])), int(rng.integers(2, 6)), int(rng.integers(5, 11))))
    _check(port, ctx, reads, _mk_pairs(rows), 11)


def test_align_golden(port, ctx, golden):
    seq, offs, k = golden["seq"], golden["offs"], int(golden["k"])
    reads = [seq[offs[i]:offs[i + 1]].tobytes() for i in range(len(offs) - 1)]
    rows = []
    for t in range(len(golden["aln_i"])):
        e = float(golden["aln_e"][t])
        rows.append((int(golden["aln_i"][t]), int(golden["aln_j"][t]), -1 if golden["aln_res"][t][5] else 1,
                     int(golden["aln_go"][t]), api.aln_limit(e, k)))
    s, q, o = _flat(reads)
    rs = ctx.upload(s, q, o)
    out = rs.align_pairs(_mk_pairs(rows), k)
    rs.close()
    for t in range(len(rows)):
        exp = golden["aln_res"][t]
        assert [out[t]["aligned"], out[t]["score"], out[t]["end_query"], out[t]["end_ref"], out[t]["comp_len"]] == exp[:5].tolist(), t


def test_align_shapes_and_ties(port, ctx):
    rng = np.random.default_rng(8)
    rnd = lambda n: bytes(ACGT[rng.integers(0, 4, n)])
    storm = lambda: bytes(np.repeat(ACGT[rng.integers(0, 4, 60)], rng.integers(1, 9, 60)))
    base = rnd(700)
    reads = [
        rnd(1), rnd(2), rnd(7), rnd(8), rnd(9), rnd(31), rnd(32), rnd(33),	       # tiny, around the 8x8 tile
        rnd(255), rnd(256), rnd(257), rnd(511), rnd(513),			       # strip boundaries (256 rows)
        base, base[100:600], base[:300] + rnd(5) + base[300:], base[50:] + rnd(80),	       # containment / indel / overhang
        storm(), storm(), storm(), b"A" * 300, b"A" * 299 + b"C", b"ACACACAC" * 40,   # tie storms
    ]
    n = len(reads)
    rows = []
    for i in range(n):
        for j in range(n):
            if (i * 7 + j * 3) % 4 == 0 or i == j:
                rows.append((i, j, 1 if (i + j) % 3 else -1, 2 + (i + j) % 4, 3 + (i * j) % 8))
    _check(port, ctx, reads, _mk_pairs(rows), 11)
    # other k (window register width) incl. the extremes the CLI allows
    _check(port, ctx, reads, _mk_pairs(rows[::5]), 10)
    _check(port, ctx, reads, _mk_pairs(rows[::7]), 31)


def test_align_non_acgt_and_case(port, ctx):
    # parasail_matrix_create("ACGT",2,-2) is case-insensitive and scores anything else 0 (memory.c:336-413);
    # the traceback marks equal characters (case-insensitively) as matches whatever they are (traceback.c match_char)
    reads = [b"ACGTNNACGTTGCANACGTACGTTGCAATGC" * 4, b"ACGTAAACGTTGCATACGTACNTTGCAATGC" * 4, b"acgtnnacgttgcanacgtacgttgcaatgc" * 4,
             b"NNNNNNNNNNNNNNNNNNNNNNNN", b"ACGTRYKMACGTTTGACCA" * 5]
    rows = [(i, j, 1, 2 + (i + j) % 4, 6) for i in range(5) for j in range(5)]
    _check(port, ctx, reads, _mk_pairs(rows), 11)


def test_align_vs_reference_long(ref, port, ctx):
    """Read-length pairs (the bench shape) against the UNMODIFIED reference incl. its 16-bit -> 32-bit path."""
    seq, _, offs = synth.make_reads(40, seed=43, reads_per_gene=8)
    reads = [seq[offs[i]:offs[i + 1]].tobytes() for i in range(40)]
    rng = np.random.default_rng(9)
    rows = [(int(rng.integers(0, 40)), int(rng.integers(0, 40)), int(rng.choice([1, -1])), int(rng.integers(2, 6)), 8)
            for _ in range(40)]
    _check(port, ctx, reads, _mk_pairs(rows), 11, ref=ref)


def test_align_empty_and_bad_args(ctx):
    s, q, o = _flat([b"ACGT" * 10, b""])
    rs = ctx.upload(s, q, o)
    assert len(rs.align_pairs(np.zeros(0, api.PAIR_DT), 11)) == 0
    with pytest.raises(api.IcbError):
        rs.align_pairs(_mk_pairs([(0, 1, 1, 2, 5)]), 11)  # empty sequence
    with pytest.raises(api.IcbError):
        rs.align_pairs(_mk_pairs([(0, 5, 1, 2, 5)]), 11)  # index outside the read set
    rs.close()


def test_align_multi_pass_and_32bit_path(port, ctx):
    """Reads above 2048 rows (several passes of the packed kernel) and above 8000 bases (32-bit kernel)."""
    rng = np.random.default_rng(10)
    rnd = lambda n: ACGT[rng.integers(0, 4, n)]

    def mutate(s, e):
        out = []
        for c in s:
            r = rng.random()
            if r < e / 3:
                continue
            if r < 2 * e / 3:
                out.append(ACGT[rng.integers(0, 4)])
            if r < e and r >= 2 * e / 3:
                c = ACGT[rng.integers(0, 4)]
            out.append(c)
        return bytes(np.array(out, np.uint8))

    g1, g2, g3 = rnd(2300), rnd(4500), rnd(8300)
    reads = [mutate(g1, 0.05), mutate(g1, 0.1), mutate(g2, 0.08), mutate(g2, 0.03), bytes(g2[500:4100]), mutate(g3, 0.06),
             mutate(g3, 0.06), bytes(rnd(2049)), bytes(rnd(2048)), bytes(rnd(4097))]
    rows = [(0, 1, 1, 3, 8), (1, 0, 1, 2, 7), (2, 3, 1, 3, 8), (3, 2, -1, 4, 9), (4, 2, 1, 3, 8), (2, 4, 1, 5, 8), (5, 6, 1, 3, 8),
            (6, 5, -1, 2, 8), (0, 2, 1, 3, 6), (7, 8, 1, 3, 5), (8, 7, 1, 3, 5), (9, 7, -1, 2, 5), (3, 9, 1, 4, 5), (5, 2, 1, 3, 8)]
    _check(port, ctx, reads, _mk_pairs(rows), 11)
