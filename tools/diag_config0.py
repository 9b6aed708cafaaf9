"""Dev tool (GPU box): where does `cluster` of the configs[0] FASTQ differ from the reference?  Runs the Python API path
against the oracle port / reference library, then both CLIs, and names the first differing section of the `.cer`."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cerfile  # noqa: E402
import oracle  # noqa: E402
from isonclust2_b200 import api, synth  # noqa: E402
from test_cluster_gpu import _gpu_cluster, _port_cluster  # noqa: E402

n, seed, rpg = int(sys.argv[1]) if len(sys.argv) > 1 else 500, int(sys.argv[2]) if len(sys.argv) > 2 else 1, 25
k, w = 11, 15
seq, qual, offs = synth.make_reads(n, seed=seed, reads_per_gene=rpg)
P = oracle.port()
p = _port_cluster(P, seq, qual, offs, k, w, 0)
ctx = api.Context(0)
for chunk in (256, 64, 1):
    g = _gpu_cluster(ctx, seq, qual, offs, k, w, 0, chunk)
    diff = np.flatnonzero((g["cluster"] != p["cluster"]) | (g["strand"] != p["strand"]))
    print(f"API chunk={chunk}: clusters {g['n_clusters']} vs port {p['n_clusters']}, aln_invoked {g['stats']['aln_invoked']} vs {p['n_aligned']}, "
          f"differing slots {len(diff)} {diff[:8]}")
    for s in diff[:4]:
        print("   slot", s, "read", g["order"][s], "gpu", g["cluster"][s], g["strand"][s], "port", p["cluster"][s], p["strand"][s])
ctx.close()

REF = os.path.join(ROOT, "oracle", "_ref", "isONclust2")
OUR = os.path.join(ROOT, "isonclust2_b200", "bin", "isONclust2")
tmp = tempfile.mkdtemp()
for who, exe in (("ref", REF), ("our", OUR)):
    d = os.path.join(tmp, who)
    os.mkdir(d)
    synth.write_fastq(os.path.join(d, "ens500.fq"), seq, qual, offs)
    for a in (["sort", "-k", "11", "-w", "15", "-o", "out", "ens500.fq"], ["cluster", "-x", "sahlin", "-l", "out/batches/isONbatch_0.cer", "-o", "clustered.cer"]):
        r = subprocess.run([exe] + a, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if r.returncode:
            print(who, a, "rc", r.returncode, r.stderr[-500:])
a, b = cerfile.parse(os.path.join(tmp, "ref/clustered.cer")), cerfile.parse(os.path.join(tmp, "our/clustered.cer"))
print("bytes equal:", a["bytes"] == b["bytes"], len(a["bytes"]), len(b["bytes"]))
for key in ("BatchNr", "BatchStart", "BatchEnd", "BatchBases", "TotalReads", "NrCls", "args", "LeftLeaf", "RightLeaf", "Depth"):
    if a[key] != b[key]:
        print("header differs:", key, a[key], b[key])
print("mindb same order:", a["mindb"] == b["mindb"], " same set:", sorted(a["mindb"]) == sorted(b["mindb"]), len(a["mindb"]), len(b["mindb"]))
print("n cls", len(a["cls"]), len(b["cls"]))
for ci, (x, y) in enumerate(zip(a["cls"], b["cls"])):
    if x != y:
        ix, iy = [m["id"] for m in x], [m["id"] for m in y]
        print("cluster", ci, "differs: sizes", len(x), len(y), "ids equal", ix == iy)
        if ix == iy:
            for m, q in zip(x, y):
                for f in m:
                    if m[f] != q[f]:
                        print("    member", m["id"], "field", f, str(m[f])[:80], "|", str(q[f])[:80])
                        break
        else:
            print("   only ref:", [i for i in ix if i not in iy][:5], "only our:", [i for i in iy if i not in ix][:5])
        break
print("graphs equal:", a["graphs"] == b["graphs"])
