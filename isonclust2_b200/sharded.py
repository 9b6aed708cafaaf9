"""MinimizerDB sharded by KEY over the ranks of a torch.distributed group, with an all-gather of every rank's partial hits.

SURVEY §8e / BASELINE.json north_star: the one place where the clustering path has a real exchange step.  The batches of a
sort and the merge jobs of one tree level shard with no collective (bench.py, tree.py); what does not shard that way is ONE
index that a single job queries — the top levels of the merge tree, where one left batch is the whole workload.  Here rank r
keeps the postings of the minimizer values v with v % world == r:

* `add` = AddMinimizers (reference src/minimizer.cpp:31-42): every rank sees every founder and adds the minimizers it owns.
* `update_cluster` = UpdateMinDB (minimizer.cpp:124-160), per shard likewise.
* `query_hits` = GetMinimizerHits (minimizer.cpp:44-76): every rank probes its shard with the read's WHOLE sketch (a key it
  does not own has no postings there), the partial hit lists are all-gathered (NCCL on device tensors, gloo on host tensors)
  and merged into the order the reference appends hits in — forward minimizers by Index, then the reverse ones; all hits of
  one minimizer come from the rank that owns its value, already in posting order, so a stable sort on (strand, Index) is the
  reference's order exactly.
* `query_counts` = the (cluster, strand) -> count / first-hit table ConsolidateMinimizerHits builds from them
  (cluster.cpp:609-615; SortMinimizerHits ranks the candidates by these counts, cluster.cpp:622-636).

The local shard is the device index behind the per-primitive C ABI (icb_index_create / _add / _update_cluster / _query_hits,
include/isonclust2_b200.h); there is no host fallback.  One query is one small launch per rank plus one collective: this is
the capacity path (an index larger than one GPU's memory), not the throughput path — the device-resident loop of
csrc/dloop.cu is, and DESIGN.md §7 says when which applies.
"""
import ctypes as C

import numpy as np

from . import api

HIT_DT = np.dtype([("cls", "<u4"), ("strand", "<i4"), ("pos", "<u4"), ("index", "<u4")])


class DeviceShard:
    """The local shard: an icb_index on this rank's GPU."""

    def __init__(self, ctx):
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        ctx.check(self.L.icb_index_create(ctx.h, C.byref(self.h)))

    def add(self, cls, mins):
        mins = np.ascontiguousarray(mins, dtype=api.MIN_DT)
        self.ctx.check(self.L.icb_index_add(self.h, int(cls), mins.ctypes.data, len(mins)))

    def update_cluster(self, cls, old_mins, new_mins):
        o = np.ascontiguousarray(old_mins, dtype=api.MIN_DT)
        n = np.ascontiguousarray(new_mins, dtype=api.MIN_DT)
        self.ctx.check(self.L.icb_index_update_cluster(self.h, int(cls), o.ctypes.data, len(o), n.ctypes.data, len(n)))

    def query_hits(self, mins, rev_mins):
        f = np.ascontiguousarray(mins, dtype=api.MIN_DT)
        r = np.ascontiguousarray(rev_mins, dtype=api.MIN_DT)
        hits = np.zeros(max(256, 4 * (len(f) + len(r))), HIT_DT)
        nh = C.c_int64()
        self.ctx.check(self.L.icb_index_query_hits(self.h, f.ctypes.data, len(f), r.ctypes.data, len(r), hits.ctypes.data, len(hits), C.byref(nh)))
        if nh.value > len(hits):
            hits = np.zeros(nh.value, HIT_DT)
            self.ctx.check(self.L.icb_index_query_hits(self.h, f.ctypes.data, len(f), r.ctypes.data, len(r), hits.ctypes.data, len(hits), C.byref(nh)))
        return hits[:nh.value]

    def close(self):
        if self.h:
            self.L.icb_index_destroy(self.h)
            self.h = C.c_void_p()


def owned(mins, rank, world):
    """The minimizers of a sketch whose VALUE this rank owns (Pos / Index keep their meaning: they belong to the cluster's
    representative, not to the shard)."""
    mins = np.asarray(mins, dtype=api.MIN_DT)
    return mins[mins["Min"] % np.uint32(world) == np.uint32(rank)]


def merge_hits(parts):
    """Partial hit lists of all ranks -> GetMinimizerHits' order (forward by Index, then reverse by Index; stable)."""
    parts = [p for p in parts if len(p)]
    if not parts:
        return np.zeros(0, HIT_DT)
    allh = np.concatenate(parts)
    key = np.where(allh["strand"] > 0, np.int64(0), np.int64(1) << 32) + allh["index"].astype(np.int64)
    return allh[np.argsort(key, kind="stable")]


def consolidate(hits):
    """ConsolidateMinimizerHits (cluster.cpp:609-615) as arrays: per (cluster, strand) the number of hits and the position of
    its first hit in the merged list (the reference's tie order starts from it), in order of first appearance."""
    if len(hits) == 0:
        z = np.zeros(0, np.int64)
        return z.astype(np.uint32), z.astype(np.int32), z, z
    key = hits["cls"].astype(np.int64) * 2 + (hits["strand"] < 0)
    uniq, first, counts = np.unique(key, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    uniq, first, counts = uniq[order], first[order], counts[order]
    return (uniq >> 1).astype(np.uint32), np.where(uniq & 1, -1, 1).astype(np.int32), counts.astype(np.int64), first.astype(np.int64)


class ShardedIndex:
    """One MinimizerDB over all ranks of `group` (default: the world).  Every rank calls every method with the same arguments
    (SPMD); every rank gets the full answer."""

    def __init__(self, ctx=None, group=None, local=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if local is None:
            if ctx is None:
                raise ValueError("ShardedIndex needs a Context (the local shard lives on its GPU)")
            local = DeviceShard(ctx)
        self.local = local
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        self.device = None
        if self.backend == "nccl":
            import torch
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.bytes_gathered = 0

    # -- updates: no communication, the key decides the owner
    def add(self, cls, mins):
        self.local.add(cls, owned(mins, self.rank, self.world))

    def update_cluster(self, cls, old_mins, new_mins):
        self.local.update_cluster(cls, owned(old_mins, self.rank, self.world), owned(new_mins, self.rank, self.world))

    # -- query: local probe + all-gather + merge
    def _allgather(self, part):
        import torch
        if self.world == 1:
            return [part]
        dist = self.dist
        dev = self.device or torch.device("cpu")
        n = torch.tensor([len(part)], dtype=torch.int64, device=dev)
        ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(ns, n, group=self.group)
        ns = [int(x.item()) for x in ns]
        cap = max(1, max(ns))
        buf = np.zeros((cap, 4), dtype=np.int32)
        buf[:len(part)] = part.view(np.int32).reshape(-1, 4)
        mine = torch.from_numpy(buf).to(dev)
        out = [torch.zeros((cap, 4), dtype=torch.int32, device=dev) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        self.bytes_gathered += cap * 16 * self.world
        return [np.ascontiguousarray(o.cpu().numpy()[:ns[r]]).view(HIT_DT).reshape(-1) for r, o in enumerate(out)]

    def query_hits(self, mins, rev_mins):
        return merge_hits(self._allgather(self.local.query_hits(mins, rev_mins)))

    def query_counts(self, mins, rev_mins):
        return consolidate(self.query_hits(mins, rev_mins))

    def close(self):
        self.local.close()
