"""isonclust2_b200 — B200-native (sm_100a) implementation of isONclust2's clustering hot path.

The product is the C-ABI shared library `libisonclust2_b200.so` (include/isonclust2_b200.h) built from
isonclust2_b200/csrc; this package is the thin host mirror used by tests and bench.py.
"""
__version__ = "0.1.0"
