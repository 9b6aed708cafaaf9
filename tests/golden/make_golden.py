#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz with the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the dev container:  python tests/golden/make_golden.py
Everything stored is an output of the reference's own functions on seeded synthetic inputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from isonclust2_b200 import synth  # noqa: E402

R = oracle.ref()
out = {}
k, w = 11, 15
N = 120
seq, qual, offs = synth.make_reads(N, seed=11, reads_per_gene=12)
# a few hand-made edge reads appended: short, homopolymer-only, exactly 2k long, w long
extra = [b"ACGT" * 3, b"A" * 50, b"ACGTTGCAAGCTTGCATGCAAG", b"ACGTACGGTCATGCA", b"AACCGGTTAACCGGTTAACCGGTTAACCGGTTAACCGGTT" * 3]
es = np.frombuffer(b"".join(extra), np.uint8)
eq = np.full(len(es), ord("5"), np.uint8)
eo = np.cumsum([0] + [len(x) for x in extra])
seq = np.concatenate([seq, es]); qual = np.concatenate([qual, eq]); offs = np.concatenate([offs, offs[-1] + eo[1:]])
N = len(offs) - 1
out.update(seq=seq, qual=qual, offs=offs, k=k, w=w)
rb = R.sort_and_sketch(seq, qual, offs, k, w)
score, err, order = rb.scores()
sk = rb.sketch()
out.update(score=score, err=err, order=order)
for key, v in sk.items():
    out["sk_" + key] = v
tab, tabn = R.qualtabs()
out.update(qualtab=tab, qualtab_nomin=tabn)
# clustering in the three modes (fresh batch each time: ClusterSortedReads mutates it)
for mode, name in ((0, "sahlin"), (1, "fast"), (2, "furious")):
    b = R.sort_and_sketch(seq, qual, offs, k, w)
    c = b.cluster(mode)
    out[f"cl_{name}_sizes"] = c["sizes"]
    out[f"cl_{name}_ids"] = c["ids"]
    out[f"cl_{name}_strands"] = c["strands"]
    out[f"cl_{name}_n_aligned"] = np.int64(c["n_aligned"])
# alignments: read vs representative pairs
rng = np.random.default_rng(5)
pa, pb, pgo, pe, res = [], [], [], [], []
for t in range(40):
    i, j = int(rng.integers(0, 120)), int(rng.integers(0, 120))
    if t % 2 == 0:
        j = i - (i % 12) + int(rng.integers(0, 12))  # same gene family more often
        j = min(j, 119)
    a = seq[offs[i]:offs[i + 1]]; b = seq[offs[j]:offs[j + 1]]
    if t % 3 == 0:
        b = R.revcomp(b)
    go = int(rng.integers(2, 6)); e = float(rng.choice([0.06, 0.1, 0.16, 0.24]))
    r = R.align(a, b, go, e, k)
    pa.append(i); pb.append(j); pgo.append(go); pe.append(e)
    res.append([r["aligned"], r["score"], r["end_query"], r["end_ref"], r["comp_len"], int(t % 3 == 0)])
out.update(aln_i=np.array(pa), aln_j=np.array(pb), aln_go=np.array(pgo), aln_e=np.array(pe), aln_res=np.array(res, np.int64))
# the reference's own unit-test goldens (test/isONclust2_test.cpp) re-derived through the library
out["ut_hpc_seq"] = R.hpc(b"AAATTTGCGTTAA", b"++:+?++++++@+")[0]
out["ut_hpc_qual"] = R.hpc(b"AAATTTGCGTTAA", b"++:+?++++++@+")[1]
out["ut_mins"] = R.minimizers(R.kmer_encode(b"ACGCCGATC", 2), 2, 4)
out["ut_err"] = np.float64(R.error_rate(b"I" * 18, nomin=False))
out["ut_pmin"] = np.float64(R.pmin_shared(13, 20, 0.111, 0.131))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
print("wrote golden_v1.npz with", len(out), "arrays;", N, "reads")
