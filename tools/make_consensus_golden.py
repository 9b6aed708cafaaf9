"""Writes tests/golden/consensus_v1.json: digests of every `.cer` the REFERENCE binary (oracle/_ref/isONclust2) writes for the
consensus-on pipeline of tests/consensus_case.py.  Run in the dev container:  python tools/make_consensus_golden.py"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import consensus_case as cc  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    cc.write_input(d)
    dig = cc.run_pipeline(os.path.join(ROOT, "oracle", "_ref", "isONclust2"), d)
with tempfile.TemporaryDirectory() as d:
    cc.write_input(d, dict(cc.P_READS))
    pdig = cc.run_pipeline(os.path.join(ROOT, "oracle", "_ref", "isONclust2"), d, steps=cc.P_STEPS, sort_args=cc.P_SORT_ARGS, n_batches=1)
out = {"sort_args": cc.SORT_ARGS, "reads": [cc.N_READS, cc.SEED, cc.READS_PER_GENE], "digests": dig,
       "period_case": {"sort_args": cc.P_SORT_ARGS, "reads": cc.P_READS, "digests": pdig}}
with open(os.path.join(ROOT, "tests", "golden", "consensus_v1.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
