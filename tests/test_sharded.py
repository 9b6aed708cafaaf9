"""Key-sharded MinimizerDB + all-gather of partial hits (isonclust2_b200/sharded.py; SURVEY §8e, north_star).

CPU (`-m "not gpu"`): the merge / consolidate logic against the oracle's sketches, and world_size 2 over gloo with a dict index
of the TEST standing in for the local shard (the product's local shard is the device index; it has no host fallback).
GPU (`-m gpu`): the same two-rank exchange with every rank's shard on the device through the C ABI, against ONE unsharded
device index and against the hit list built from the oracle's sketches.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from isonclust2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class HostShard:
    """Test stand-in for a local shard: MinimizerDB as the reference keeps it (minimizer.h:60-61), AddMinimizers /
    GetMinimizerHits semantics (minimizer.cpp:31-76)."""

    def __init__(self):
        self.db = {}

    def add(self, cls, mins):
        for m in mins:  # a cluster enters a posting list once (minimizer.cpp:31-42: only if cls > back())
            lst = self.db.setdefault(int(m["Min"]), [])
            if not lst or int(cls) > lst[-1]:
                lst.append(int(cls))

    def query_hits(self, mins, rev_mins):
        from isonclust2_b200.sharded import HIT_DT
        out = []
        for strand, arr in ((1, mins), (-1, rev_mins)):
            for m in arr:
                for c in self.db.get(int(m["Min"]), ()):
                    out.append((c, strand, int(m["Pos"]), int(m["Index"])))
        return np.array(out, dtype=HIT_DT) if out else np.zeros(0, HIT_DT)

    def close(self):
        pass


def _sketches(port, n=40, seed=77):
    seq, qual, offs = synth.make_reads(n, seed=seed, reads_per_gene=5, min_len=300, max_len=700)
    sk = port.sketch_batch(seq, qual, offs, 11, 15)
    good = [i for i in range(n) if sk["status"][i] == 1]
    f = lambda i: np.ascontiguousarray(sk["mins"][sk["min_offs"][i]:sk["min_offs"][i + 1]])
    r = lambda i: np.ascontiguousarray(sk["rev_mins"][sk["rev_offs"][i]:sk["rev_offs"][i + 1]])
    return good, f, r


def test_merge_of_key_shards_is_the_unsharded_hit_list(port):
    """Any number of shards, merged: exactly GetMinimizerHits' list (order included) of one index."""
    from isonclust2_b200 import sharded
    good, f, r = _sketches(port)
    reps, queries = good[:15], good[15:30]
    whole = HostShard()
    for c, i in enumerate(reps):
        whole.add(c, f(i))
    for world in (1, 2, 3, 8):
        shards = [HostShard() for _ in range(world)]
        for c, i in enumerate(reps):
            parts = [sharded.owned(f(i), rk, world) for rk in range(world)]
            assert sum(len(p) for p in parts) == len(f(i))
            for rk in range(world):
                shards[rk].add(c, parts[rk])
        for q in queries:
            want = whole.query_hits(f(q), r(q))
            got = sharded.merge_hits([s.query_hits(f(q), r(q)) for s in shards])
            assert got.tolist() == want.tolist()
            cls, strand, cnt, first = sharded.consolidate(got)
            # ConsolidateMinimizerHits: one entry per (cluster, strand), sizes add up, first-appearance order
            assert int(cnt.sum()) == len(want) and len(set(zip(cls.tolist(), strand.tolist()))) == len(cls)
            assert (np.diff(first) > 0).all()
            for c, s, n in zip(cls, strand, cnt):
                assert n == int(((want["cls"] == c) & (want["strand"] == s)).sum())


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import torch.distributed as dist
dist.init_process_group({backend!r})
rank, world = dist.get_rank(), dist.get_world_size()
import oracle
from isonclust2_b200 import sharded
import test_sharded as T
port = oracle.port()
good, f, r = T._sketches(port)
reps, queries = good[:15], good[15:30]
device = {device!r}
ctx = None
if device:
    from isonclust2_b200 import api
    if {backend!r} == "nccl":
        import torch
        torch.cuda.set_device(rank)
    ctx = api.Context(rank if {backend!r} == "nccl" else 0)
    idx = sharded.ShardedIndex(ctx)
    whole = sharded.DeviceShard(ctx)
else:
    idx = sharded.ShardedIndex(local=T.HostShard())
    whole = T.HostShard()
assert idx.world == world == 2
for c, i in enumerate(reps):
    idx.add(c, f(i))
    whole.add(c, f(i))
truth = T.HostShard()
for c, i in enumerate(reps):
    truth.add(c, f(i))
nhits = 0
for q in queries:
    got = idx.query_hits(f(q), r(q))
    assert got.tolist() == whole.query_hits(f(q), r(q)).tolist() == truth.query_hits(f(q), r(q)).tolist(), q
    nhits += len(got)
    part = idx.local.query_hits(f(q), r(q))  # this rank's shard only holds its own keys
    mins = {{int(m["Index"]): int(m["Min"]) for m in f(q)}}
    rmins = {{int(m["Index"]): int(m["Min"]) for m in r(q)}}
    for h in part:
        v = (mins if h["strand"] > 0 else rmins)[int(h["index"])]
        assert v % world == rank
assert nhits > 0 and idx.bytes_gathered > 0
idx.close(); whole.close()
if ctx: ctx.close()
dist.barrier()
if rank == 0:
    print("SHARDED OK", nhits)
dist.destroy_process_group()
"""


def _run_two_ranks(tmp_path, backend, device, port_no):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, backend=backend, device=device))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port_no), str(script)], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "SHARDED OK" in r.stdout


def test_two_ranks_gloo_host_shards(tmp_path):
    """world_size 2 over gloo on CPU: ownership by key, all-gather of ragged partial lists, merge == one index."""
    _run_two_ranks(tmp_path, "gloo", False, 29641)


@pytest.mark.gpu
def test_two_ranks_device_shards(tmp_path):
    """Both ranks keep their shard in a device index (C ABI) — NCCL with one GPU per rank where the box has two, else gloo with
    both ranks on GPU 0 — and the merged hits equal those of one unsharded device index and of the oracle's sketches."""
    import torch
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    _run_two_ranks(tmp_path, backend, True, 29642)


@pytest.mark.gpu
def test_single_rank_device_index(port, ctx):
    """world 1 (no process group): the sharded index degenerates to the device index."""
    from isonclust2_b200 import sharded
    good, f, r = _sketches(port)
    idx = sharded.ShardedIndex(ctx)
    truth = HostShard()
    for c, i in enumerate(good[:15]):
        idx.add(c, f(i))
        truth.add(c, f(i))
    for q in good[15:30]:
        assert idx.query_hits(f(q), r(q)).tolist() == truth.query_hits(f(q), r(q)).tolist()
    idx.close()
