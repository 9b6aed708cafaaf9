#!/usr/bin/env python
"""Dev tool: cluster N synthetic reads with the product (GPU) and with the unmodified reference, report differences."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import oracle
from isonclust2_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
k, w = 11, 15
seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,))
R = oracle.ref()
t = time.time(); rb = R.sort_and_sketch(seq, qual, offs, k, w); _, _, order = rb.scores(); rc = rb.cluster(mode); tr = time.time() - t
ctx = api.Context(0)
t = time.time(); g = api.run_pipeline(ctx, seq, qual, offs, k=k, w=w, mode=mode, batch_kb=0, batch_max_seq=0); tg = time.time() - t
print("ref clusters", rc["n_clusters"], "aligned", rc["n_aligned"], "%.1fs" % tr, "| gpu clusters", g["n_clusters"], g["stats"], "%.1fs" % tg)
ra, pos = {}, 0
for c, sz in enumerate(rc["sizes"]):
    for j in range(1, sz):
        ra[int(rc["ids"][pos + j])] = (c, int(rc["strands"][pos + j]))
    pos += sz
ga = {int(g["order"][s]): (int(g["cluster"][s]), int(g["strand"][s])) for s in range(n) if g["cluster"][s] >= 0}
bad = [r for r in ra if ga.get(r) != ra[r]]
print("reads", len(ra), len(ga), "mismatching", len(bad))
slot = {int(r): s for s, r in enumerate(g["order"])}
for r in sorted(bad, key=lambda r: slot[r])[:10]:
    print(" slot", slot[r], "read", r, "ref", ra[r], "gpu", ga.get(r))
