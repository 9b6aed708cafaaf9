"""Host partial-order alignment (csrc/poa.cpp, icb_poa_*) against the reference's vendored spoa: bit-exact alignments,
consensus sequences and serialized graphs (the ConsGs blobs of a `.cer`).  CPU only — the consensus stays on the host.

  golden    tests/golden/poa_v1.json, written by the reference (tools/make_poa_golden.py)
  live      oracle/_ref (the unmodified reference compiled here), when present: the three alignment types, the int16 and
            the int32 engine, ragged reads, weights, reloaded graphs, Clear + re-add (ConsPurge, consensus.cpp:128-137)
"""
import json
import os

import numpy as np
import pytest

import poa_util as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    return pu.ours()


@pytest.fixture(scope="module")
def R(ref):  # noqa: ARG001 — skips when oracle/_ref is absent
    return pu.spoa()


def test_poa_golden(L):
    with open(os.path.join(ROOT, "tests", "golden", "poa_v1.json")) as f:
        gold = json.load(f)
    assert len(gold["cases"]) == len(pu.GOLDEN_CASES)
    for g in gold["cases"]:
        assert pu.digest_case(pu.OurGraph(L), tuple(g["case"])) == g["sha256"], g["case"]


def _lockstep(L, R, reads, algo, check_align=True):
    og, rg = pu.OurGraph(L), pu.SpoaGraph(R)
    for t, (s, w) in enumerate(reads):
        if check_align:
            a1, b1, s1 = og.align(s, algo)
            a2, b2, s2 = rg.align(s, algo)
            assert np.array_equal(a1, a2) and np.array_equal(b1, b2), (t, "alignment")
            if len(a1):
                assert s1 == s2, (t, "score")
        og.add(s, w, algo)
        rg.add(s, w, algo)
        assert og.consensus() == rg.consensus(), (t, "consensus")
        assert og.save() == rg.save(), (t, "graph bytes")
    return og, rg


@pytest.mark.parametrize("algo", [2, 0, 1])
@pytest.mark.parametrize("seed,length,n,trunc", [(1, 300, 10, False), (2, 450, 10, True), (3, 700, 8, True)])
def test_poa_vs_spoa(L, R, algo, seed, length, n, trunc):
    _lockstep(L, R, pu.family(seed, length, n, trunc=trunc), algo)


def test_poa_int32_engine(L, R):
    # worst-case score below int16: the engine switches to 8 lanes of int32 (simd_alignment_engine_implementation.hpp:718-748)
    _lockstep(L, R, pu.family(7, 4200, 4), 2)
    # and a family whose graph crosses the bound while it grows (gap(len + 8) + gap(nodes) < -31744: len + 8 + nodes > 7936)
    reads = pu.family(8, 3700, 8)
    og, _ = _lockstep(L, R, reads, 2)
    assert len(reads[1][0]) + 8 + len(reads[0][0]) < 7936 < len(reads[-1][0]) + 8 + og.sizes()[1]


def test_poa_edge_cases(L, R):
    og, rg = pu.OurGraph(L), pu.SpoaGraph(R)
    a, b, _ = og.align(b"ACGT", 2)  # empty graph: empty alignment
    assert len(a) == 0 and len(b) == 0
    for s, w in ((b"A", 1), (b"ACGTTGCA", 2), (b"TTTT", 1), (b"ACGTTGCAACGTTGCAAC", 3), (b"G", 5), (b"ACGNNTGCA", 1)):
        og.add(s, w, 2)
        rg.add(s, w, 2)
        assert og.save() == rg.save(), s
        assert og.consensus() == rg.consensus(), s
    og.add(b"", 1, 2)  # AddAlignment returns at once for an empty sequence
    rg.add(b"", 1, 2)
    assert og.save() == rg.save()


def test_poa_reload_and_purge(L, R):
    reads = pu.family(21, 500, 12, trunc=True)
    og, rg = _lockstep(L, R, reads[:6], 2, check_align=False)
    blob = og.save()
    # a graph read back from a `.cer` has its edge lists in file order: both sides continue from the same bytes
    og2, rg2 = pu.OurGraph(L, blob), pu.SpoaGraph(R, blob)
    assert og2.save() == blob == rg2.save()
    for s, w in reads[6:]:
        og2.add(s, w, 2)
        rg2.add(s, w, 2)
        assert og2.save() == rg2.save()
        assert og2.consensus() == rg2.consensus()
    # ConsPurge: Clear, then the consensus re-added with the number of sequences as its weight
    cons, nseq = og2.consensus(), og2.sizes()[0]
    og2.clear()
    rg2.clear()
    og2.add(cons, nseq, 2)
    rg2.add(cons, nseq, 2)
    assert og2.save() == rg2.save()
    assert og2.sizes()[0] == 1


def test_poa_load_rejects_garbage(L):
    import ctypes as C
    h = C.c_void_p()
    assert L.icb_poa_graph_load(b"\x01\x02\x03", 3, C.byref(h)) != 0
    good = pu.OurGraph(L)
    good.add(b"ACGTACGT", 1, 2)
    blob = good.save()
    assert L.icb_poa_graph_load(blob[:-3], len(blob) - 3, C.byref(h)) != 0
    assert L.icb_poa_graph_load(blob + b"\0", len(blob) + 1, C.byref(h)) != 0
    # a stored order that is not a forward permutation of the nodes (spoa would walk outside its matrices)
    import struct
    n_nodes = good.sizes()[1]
    tail = 8 + 4 * n_nodes + 8  # rank_to_node_id (count + ids) + the empty consensus vector
    ranks = list(struct.unpack_from(f"<{n_nodes}I", blob, len(blob) - tail + 8))
    assert sorted(ranks) == list(range(n_nodes))
    for bad in ([ranks[0]] * n_nodes, ranks[::-1]):
        b2 = bytearray(blob)
        struct.pack_into(f"<{n_nodes}I", b2, len(blob) - tail + 8, *bad)
        assert L.icb_poa_graph_load(bytes(b2), len(b2), C.byref(h)) != 0
    b3 = bytearray(blob)
    struct.pack_into("<I", b3, 4 + 2 * 1032 + 8 + 1 + 4, 77)  # first node's code
    assert L.icb_poa_graph_load(bytes(b3), len(b3), C.byref(h)) != 0


def test_python_mirror(L):
    """api.PoaGraph (the ctypes mirror the Python host side uses) over the same entry points"""
    from isonclust2_b200 import api
    reads = pu.family(31, 300, 6)
    g, og = api.PoaGraph(), pu.OurGraph(L)
    try:
        for s, w in reads:
            g.add(s, w, 2)
            og.add(s, w, 2)
        assert g.n_sequences == 6 and g.consensus() == og.consensus() and g.save() == og.save()
        g2 = api.PoaGraph(g.save())
        assert g2.save() == g.save()
        g2.clear()
        assert g2.n_sequences == 0
        g2.close()
        with pytest.raises(api.IcbError):
            api.PoaGraph(b"not a graph")
    finally:
        g.close()
