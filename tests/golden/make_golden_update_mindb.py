#!/usr/bin/env python
"""Generate tests/golden/golden_update_mindb.npz with the UNMODIFIED reference (oracle/_ref): a MinDB built by
AddMinimizers (minimizer.cpp:31-42) and patched by UpdateMinDB (minimizer.cpp:124-160), exported sorted by key.

Run in the dev container:  python tests/golden/make_golden_update_mindb.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from test_oracle import _update_case  # noqa: E402  (the same case the live comparison uses)

case, _ = _update_case(oracle.port())
cm, co, uc, um, uo = case
keys, offs, cls = oracle.ref().mindb_after_updates(cm, co, uc, um, uo)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_update_mindb.npz"), cls_mins=cm, cls_offs=co,
                    upd_cls=np.asarray(uc, np.int32), upd_mins=um, upd_offs=uo, keys=keys, offs=offs, cls=cls)
print("keys", len(keys), "postings", len(cls), "empty lists", int((offs[1:] == offs[:-1]).sum()))
