"""Empirical shared-minimizer probability table (reference: src/p_emp_prob.cpp:22-94).

`load_p_shared(k, w)` returns the 15x15 float64 grid p_shared[c1, c2] for error
classes c = round(e*100)-1 (0.01 .. 0.15) that InitMinSharedMap would build for
(k, w): rows with the same k and |w_row - w| <= 2, later rows overwriting
earlier ones, symmetric in (e1, e2).  Cells the table does not cover are NaN
(the reference throws std::invalid_argument on lookup, p_emp_prob.cpp:88).
"""
import os
import struct

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "p_emp_prob.bin")
_rows = None


def _load():
    global _rows
    if _rows is None:
        with open(_DATA, "rb") as f:
            assert f.read(8) == b"ICBPEMP1"
            (n,) = struct.unpack("<I", f.read(4))
            _rows = np.frombuffer(f.read(), dtype=np.dtype([("k", "u1"), ("w", "u1"), ("c1", "u1"), ("c2", "u1"), ("p", "<f8")]), count=n)
    return _rows


def load_p_shared(k: int, w: int) -> np.ndarray:
    rows = _load()
    sel = rows[(rows["k"] == k) & (np.abs(rows["w"].astype(np.int64) - w) <= 2)]
    tab = np.full((15, 15), np.nan, dtype=np.float64)
    for r in sel:  # file order: later rows overwrite
        tab[r["c1"] - 1, r["c2"] - 1] = r["p"]
        tab[r["c2"] - 1, r["c1"] - 1] = r["p"]
    return tab
