"""Batch / merge-tree scheduler over the drop-in CLI: what the reference's Snakemake pipeline does around the binary
(README.md:105-117 — `sort`, one `cluster` job per batch, then pairwise merges `cluster -l A -r B` level by level), with
the jobs of one level sharded across the ranks of a `torch.distributed` group: one process per GPU, no data-path
collective — a barrier between levels is the only synchronisation, files live on the node's file system.

The tree keeps the reference's constraints (cluster.cpp:80-90): a merged right batch must start where the left one ends,
and the left batch must not be shallower than the right one.  Adjacent nodes are paired left to right; an odd last node
is carried up unchanged, so it only ever becomes a RIGHT child.

    python -m isonclust2_b200.tree --fastq reads.fq --out work [--mode sahlin] [-- sort flags ...]
    torchrun --nproc-per-node N -m isonclust2_b200.tree ...        (rank r drives GPU r)
"""
import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CLI = os.path.join(HERE, "bin", "isONclust2")


class Node:
    """One `.cer` of the tree: batches [lo, hi], `depth` as the reference counts it (0 = clustered leaf)."""

    def __init__(self, lo, hi, depth, path, left=None, right=None):
        self.lo, self.hi, self.depth, self.path, self.left, self.right = lo, hi, depth, path, left, right

    def __repr__(self):
        return f"Node({self.lo}-{self.hi} d{self.depth} {self.path})"


def plan(n_batches, workdir="."):
    """levels[0] = the leaf jobs (cluster of one sorted batch), levels[l>0] = merge jobs; returns (levels, root).

    Every job of a level only needs outputs of earlier levels, so the jobs of one level are independent."""
    if n_batches <= 0:
        return [], None
    nodes = [Node(i, i, 0, os.path.join(workdir, f"c_{i}.cer")) for i in range(n_batches)]
    levels = [list(nodes)]
    while len(nodes) > 1:
        nxt, jobs = [], []
        for i in range(0, len(nodes) - 1, 2):
            a, b = nodes[i], nodes[i + 1]
            assert b.lo == a.hi + 1 and a.depth >= b.depth, (a, b)
            m = Node(a.lo, b.hi, a.depth + 1, os.path.join(workdir, f"m_{a.lo}_{b.hi}.cer"), a, b)
            nxt.append(m)
            jobs.append(m)
        if len(nodes) % 2:
            nxt.append(nodes[-1])
        levels.append(jobs)
        nodes = nxt
    return levels, nodes[0]


def shard(jobs, rank, world):
    """Jobs of one level owned by `rank`: round-robin, like batches -> GPUs in SURVEY §8e."""
    return [j for i, j in enumerate(jobs) if i % world == rank]


def job_command(node, batch_dir, mode, cli=CLI, extra=()):
    if node.left is None:
        return [cli, "cluster", "-x", mode, "-l", os.path.join(batch_dir, f"isONbatch_{node.lo}.cer"), "-o", node.path, *extra]
    return [cli, "cluster", "-x", mode, "-l", node.left.path, "-r", node.right.path, "-o", node.path, *extra]


def run_tree(n_batches, batch_dir, workdir, mode="sahlin", rank=0, world=1, barrier=None, runner=None, cli=CLI, device=None, extra=(), agree=None,
             cwd=None):
    """Runs this rank's share of every level; `barrier()` separates levels (or `agree(ok)`, which also tells every rank
    whether all of them succeeded); `runner(cmd, env)` executes one job (default: subprocess with ICB_DEVICE set to this
    rank's GPU) in directory `cwd` (where relative job outputs live; default: the current one).  Returns the root node."""
    levels, root = plan(n_batches, workdir)
    env = dict(os.environ)
    if device is not None:
        env["ICB_DEVICE"] = str(device)
    if world > 1 and "ICB_HOST_CORES" not in env:
        # the ranks of one node share its cores: each job's helper threads are sized for this rank's share
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        env["ICB_HOST_CORES"] = str(max(1, cores // max(local_world, 1)))

    def default_runner(cmd, env):
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{' '.join(cmd)} failed ({r.returncode}): {r.stderr[-2000:]}")

    runner = runner or default_runner
    for jobs in levels:
        # A failed job must not leave the other ranks waiting in the barrier, nor a half-written output for the next level:
        # every job writes to a temporary name that is renamed on success, and the ranks agree on the level's outcome
        # (`agree(ok) -> bool`, an all-reduce of a status flag) in place of a bare barrier.
        error = None
        for node in shard(jobs, rank, world):
            final = node.path
            node.path = final + ".partial"
            try:
                runner(job_command(node, batch_dir, mode, cli, extra), env)
                os.replace(os.path.join(cwd or "", node.path), os.path.join(cwd or "", final))
            except BaseException as e:  # noqa: BLE001 - reported on every rank below
                error = e
                break
            finally:
                node.path = final
        if agree is not None:
            if not agree(error is None):
                raise RuntimeError(f"merge tree: a job of this level failed on {'this' if error else 'another'} rank") from error
        else:
            if error is not None:
                raise error
            if barrier is not None:
                barrier()
    return root


def make_agree(dist, device):
    """`agree(ok)`: True on every rank iff every rank passed True — the barrier between tree levels that also carries the
    outcome (one MIN all-reduce of a flag), so that no rank waits for a rank that has already given up."""
    import torch

    def agree(ok):
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())
    return agree


def count_batches(batch_dir):
    n = 0
    while os.path.exists(os.path.join(batch_dir, f"isONbatch_{n}.cer")):
        n += 1
    return n


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--fastq", required=True)
    ap.add_argument("--out", required=True, help="work directory (sorted batches, clustered and merged .cer files)")
    ap.add_argument("--mode", default="sahlin", choices=["sahlin", "fast", "furious"])
    ap.add_argument("--backend", default=None, help="torch.distributed backend when WORLD_SIZE > 1 (default: nccl with a GPU, else gloo)")
    ap.add_argument("sort_flags", nargs="*", help="extra flags for `isONclust2 sort` (after --)")
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    agree = None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        backend = args.backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend)
        agree = make_agree(dist, torch.device("cuda", local) if backend == "nccl" else torch.device("cpu"))
    sort_dir = os.path.join(args.out, "sorted")
    ok = True
    if rank == 0:
        os.makedirs(args.out, exist_ok=True)
        env = dict(os.environ, ICB_DEVICE=str(local))
        r = subprocess.run([CLI, "sort", "-o", sort_dir, *args.sort_flags, args.fastq], env=env, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            ok = False
    if agree:
        ok = agree(ok)
    if not ok:
        if dist is not None:
            dist.destroy_process_group()
        raise SystemExit(1)
    batch_dir = os.path.join(sort_dir, "batches")
    root = run_tree(count_batches(batch_dir), batch_dir, args.out, args.mode, rank, world, device=local, agree=agree)
    if rank == 0:
        print(root.path if root else "")
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
