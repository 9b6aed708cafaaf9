"""Synthetic ONT-like cDNA reads (ctypes front end of tools/synth.c).  Data tooling."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libicb_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = ctypes.CDLL(path)
        _lib.synth_sizes.restype = ctypes.c_int64
        _lib.synth_sizes.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.synth_fill_g.restype = None
        _lib.synth_sizes_g.restype = ctypes.c_int64
        _lib.synth_sizes_g.argtypes = _lib.synth_sizes.argtypes
        _lib.synth_fill_g.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.synth_fill.restype = None
        _lib.synth_fill.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def make_reads(n_reads, seed=1, reads_per_gene=25, min_len=600, max_len=2000, err_set=(0.03, 0.05, 0.08, 0.12),
               seq_out=None, qual_out=None, n_genes=None):
    """Returns (seq uint8[total], qual uint8[total], offs int64[n+1]).

    n_genes: gene count (default n_reads // reads_per_gene); read i is a function of (seed, i, genes), so
    make_reads(n, n_genes=N // 25) is the first n reads of make_reads(N).
    seq_out/qual_out: optional callables size -> uint8 ndarray (e.g. pinned host memory)."""
    lib = _load()
    errs = np.ascontiguousarray(err_set, dtype=np.float64)
    offs = np.zeros(n_reads + 1, dtype=np.int64)
    ng = int(n_genes) if n_genes is not None else max(1, n_reads // reads_per_gene)
    total = lib.synth_sizes_g(seed, n_reads, ng, min_len, max_len, errs.ctypes.data, len(errs), offs.ctypes.data)
    seq = seq_out(total) if seq_out else np.empty(total, dtype=np.uint8)
    qual = qual_out(total) if qual_out else np.empty(total, dtype=np.uint8)
    lib.synth_fill_g(seed, n_reads, ng, min_len, max_len, errs.ctypes.data, len(errs), offs.ctypes.data,
                     seq.ctypes.data, qual.ctypes.data)
    return seq, qual, offs


def write_fastq(path, seq, qual, offs):
    with open(path, "wb") as f:
        for i in range(len(offs) - 1):
            a, b = int(offs[i]), int(offs[i + 1])
            f.write(b"@read_%d\n" % i + seq[a:b].tobytes() + b"\n+\n" + qual[a:b].tobytes() + b"\n")
