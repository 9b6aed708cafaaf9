"""Writes tests/golden/poa_v1.json: digests of what the reference's vendored spoa (through oracle/_ref) produces on the seeded
read families of tests/poa_util.py — alignment, consensus and serialized graph after every added read.  Run in the dev
container (needs oracle/_ref):  python tools/make_poa_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import poa_util as pu  # noqa: E402

R = pu.spoa()
out = {"cases": [{"case": list(c), "sha256": pu.digest_case(pu.SpoaGraph(R), c)} for c in pu.GOLDEN_CASES]}
with open(os.path.join(ROOT, "tests", "golden", "poa_v1.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
