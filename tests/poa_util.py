"""ctypes faces of the host POA (icb_poa_*, the product) and of spoa through the reference shim (ref_poa_*, the checker),
plus the seeded read families both are fed with.  Test infrastructure."""
import ctypes as C
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_p, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int


def ours():
    # ICB_POA_LIB: another build of csrc/poa.cpp (e.g. one without the AVX2 clone) under the same tests
    L = C.CDLL(os.environ.get("ICB_POA_LIB") or os.path.join(ROOT, "isonclust2_b200", "libisonclust2_b200.so"))
    L.icb_poa_graph_create.argtypes = [C.POINTER(_p)]
    L.icb_poa_graph_destroy.argtypes = [_p]
    L.icb_poa_graph_destroy.restype = None
    L.icb_poa_graph_load.argtypes = [_p, _i64, C.POINTER(_p)]
    L.icb_poa_graph_add.argtypes = [_p, C.c_char_p, _i64, C.c_uint, _i32]
    L.icb_poa_align.argtypes = [_p, C.c_char_p, _i64, _i32, _p, _p, _i64, C.POINTER(_i64), C.POINTER(C.c_int32)]
    L.icb_poa_graph_consensus.argtypes = [_p, _p, _i64, C.POINTER(_i64)]
    L.icb_poa_graph_save.argtypes = [_p, _p, _i64, C.POINTER(_i64)]
    L.icb_poa_graph_sizes.argtypes = [_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]
    L.icb_poa_graph_clear.argtypes = [_p]
    return L


def spoa():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libisonclust2_ref.so"))
    R.ref_poa_new.restype = _p
    R.ref_poa_free.argtypes = [_p]
    R.ref_poa_free.restype = None
    R.ref_poa_clear.argtypes = [_p]
    R.ref_poa_clear.restype = None
    R.ref_poa_add.argtypes = [_p, C.c_char_p, _i32, C.c_uint, _i32]
    R.ref_poa_add.restype = None
    R.ref_poa_align.restype = _i64
    R.ref_poa_align.argtypes = [_p, C.c_char_p, _i32, _i32, _p, _p, _i64, _p]
    R.ref_poa_consensus.restype = _i64
    R.ref_poa_consensus.argtypes = [_p, _p, _i64]
    R.ref_poa_save.restype = _i64
    R.ref_poa_save.argtypes = [_p, _p, _i64]
    R.ref_poa_load.restype = _p
    R.ref_poa_load.argtypes = [_p, _i64]
    R.ref_poa_num_sequences.restype = _i64
    R.ref_poa_num_sequences.argtypes = [_p]
    return R


def mutate(rng, s, err):
    out = bytearray()
    for ch in s:
        r = rng.random()
        if r < err / 3:
            continue
        if r < 2 * err / 3:
            out.append(int(rng.choice(list(b"ACGT"))))
            out.append(ch)
            continue
        if r < err:
            out.append(int(rng.choice(list(b"ACGT"))))
            continue
        out.append(ch)
    return bytes(out)


def family(seed, length, n, err=0.08, trunc=False):
    """n noisy copies of one random transcript (optionally with ragged ends) and a weight 1..3 for each"""
    rng = np.random.default_rng(seed)
    base = bytes(rng.choice(list(b"ACGT"), length).tolist())
    out = []
    for _ in range(n):
        s = mutate(rng, base, err)
        if trunc:
            a = int(rng.integers(0, max(1, len(s) // 5)))
            b = int(rng.integers(0, max(1, len(s) // 5)))
            s = s[a:len(s) - b]
        out.append((s, int(rng.integers(1, 4))))
    return out


class OurGraph:
    def __init__(self, L, blob=None):
        self.L, self.h = L, _p()
        if blob is None:
            assert L.icb_poa_graph_create(C.byref(self.h)) == 0
        else:
            assert L.icb_poa_graph_load(blob, len(blob), C.byref(self.h)) == 0

    def __del__(self):
        self.L.icb_poa_graph_destroy(self.h)

    def align(self, s, algo):
        n, sc = _i64(0), C.c_int32(0)
        cap = len(s) + 400000
        a, b = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        assert self.L.icb_poa_align(self.h, s, len(s), algo, a.ctypes.data, b.ctypes.data, cap, C.byref(n), C.byref(sc)) == 0
        assert n.value <= cap
        return a[:n.value].copy(), b[:n.value].copy(), sc.value

    def add(self, s, w, algo):
        assert self.L.icb_poa_graph_add(self.h, s, len(s), w, algo) == 0

    def consensus(self):
        n = _i64(0)
        assert self.L.icb_poa_graph_consensus(self.h, None, 0, C.byref(n)) == 0
        buf = C.create_string_buffer(max(1, n.value))
        assert self.L.icb_poa_graph_consensus(self.h, buf, n.value, C.byref(n)) == 0
        return buf.raw[:n.value]

    def save(self):
        n = _i64(0)
        assert self.L.icb_poa_graph_save(self.h, None, 0, C.byref(n)) == 0
        buf = C.create_string_buffer(max(1, n.value))
        assert self.L.icb_poa_graph_save(self.h, buf, n.value, C.byref(n)) == 0
        return buf.raw[:n.value]

    def sizes(self):
        a, b, c = _i64(0), _i64(0), _i64(0)
        assert self.L.icb_poa_graph_sizes(self.h, C.byref(a), C.byref(b), C.byref(c)) == 0
        return a.value, b.value, c.value

    def clear(self):
        assert self.L.icb_poa_graph_clear(self.h) == 0


class SpoaGraph:
    CAP = 1 << 26

    def __init__(self, R, blob=None):
        self.R = R
        self.h = R.ref_poa_new() if blob is None else R.ref_poa_load(blob, len(blob))

    def __del__(self):
        self.R.ref_poa_free(self.h)

    def align(self, s, algo):
        cap = len(s) + 400000
        a, b, sc = np.zeros(cap, np.int32), np.zeros(cap, np.int32), C.c_int32(0)
        n = self.R.ref_poa_align(self.h, s, len(s), algo, a.ctypes.data, b.ctypes.data, cap, C.byref(sc))
        assert n <= cap
        return a[:n].copy(), b[:n].copy(), sc.value

    def add(self, s, w, algo):
        self.R.ref_poa_add(self.h, s, len(s), w, algo)

    def consensus(self):
        buf = C.create_string_buffer(self.CAP)
        n = self.R.ref_poa_consensus(self.h, buf, self.CAP)
        assert n <= self.CAP
        return buf.raw[:n]

    def save(self):
        buf = C.create_string_buffer(self.CAP)
        n = self.R.ref_poa_save(self.h, buf, self.CAP)
        assert n <= self.CAP
        return buf.raw[:n]

    def clear(self):
        self.R.ref_poa_clear(self.h)


GOLDEN_CASES = [  # (seed, transcript length, reads, ragged ends, algo)
    (101, 400, 10, False, 2), (102, 650, 14, True, 2), (103, 500, 8, True, 0), (104, 450, 8, False, 1), (105, 1300, 30, True, 2),
]


def digest_case(graph, case):
    """sha256 over, per added read: the alignment, the consensus and the serialized graph"""
    seed, length, n, trunc, algo = case
    h = hashlib.sha256()
    for s, w in family(seed, length, n, trunc=trunc):
        a, b, _ = graph.align(s, algo)
        h.update(a.tobytes())
        h.update(b.tobytes())
        graph.add(s, w, algo)
        h.update(graph.consensus())
        h.update(graph.save())
    return h.hexdigest()
