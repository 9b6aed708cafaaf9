"""Merge-tree scheduler (isonclust2_b200/tree.py): plan properties, world_size-2 sharding over gloo on CPU (with the
reference binary as the job executable, so the reference's own consistency checks — consecutive ranges, left depth >=
right depth, cluster.cpp:80-90 — police the plan), and on the GPU the same tree through the drop-in CLI."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "isONclust2")
OUR_BIN = os.path.join(ROOT, "isonclust2_b200", "bin", "isONclust2")


def test_plan_properties():
    from isonclust2_b200 import tree
    assert tree.plan(0) == ([], None)
    for n in range(1, 70):
        levels, root = tree.plan(n, "w")
        assert (root.lo, root.hi) == (0, n - 1)
        done = set()
        for li, jobs in enumerate(levels):
            for j in jobs:
                if li == 0:
                    assert j.left is None and j.lo == j.hi and j.depth == 0
                else:
                    a, b = j.left, j.right
                    assert a.path in done and b.path in done  # inputs come from earlier levels only
                    assert b.lo == a.hi + 1 and a.depth >= b.depth and j.depth == a.depth + 1
                    assert (j.lo, j.hi) == (a.lo, b.hi)
            done |= {j.path for j in jobs}
        assert len(levels[0]) == n and sum(len(l) for l in levels[1:]) == n - 1
        # round-robin sharding partitions every level
        for world in (1, 2, 3, 8):
            for jobs in levels:
                parts = [tree.shard(jobs, r, world) for r in range(world)]
                assert sorted(j.path for p in parts for j in p) == sorted(j.path for j in jobs)


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from isonclust2_b200 import tree
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
log = []
def runner(cmd, env):
    import subprocess
    log.append(" ".join(cmd[1:]))
    r = subprocess.run(cmd, cwd={cwd!r}, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, (cmd, r.stderr[-1500:])
import torch
root = tree.run_tree({n}, "srt/batches", "par", "sahlin", rank, world, runner=runner, cli={cli!r}, cwd={cwd!r},
                     agree=tree.make_agree(dist, torch.device("cpu")))
gathered = [None] * world
dist.all_gather_object(gathered, log)
if rank == 0:
    assert all(len(g) > 0 for g in gathered), gathered
    assert sum(len(g) for g in gathered) == 2 * {n} - 1
    print("ROOT", root.path)
dist.destroy_process_group()
"""


def _sorted_batches(tmp_path, exe, n_reads=1000, per_batch=200):
    from isonclust2_b200 import synth
    seq, qual, offs = synth.make_reads(n_reads, seed=11)
    synth.write_fastq(str(tmp_path / "reads.fq"), seq, qual, offs)
    r = subprocess.run([exe, "sort", "-M", str(per_batch), "-o", "srt", "reads.fq"], cwd=tmp_path, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    from isonclust2_b200 import tree
    return tree.count_batches(str(tmp_path / "srt" / "batches"))


def _serial_tree(tmp_path, exe, n, outdir):
    from isonclust2_b200 import tree
    os.makedirs(tmp_path / outdir, exist_ok=True)

    def runner(cmd, env):
        r = subprocess.run(cmd, cwd=tmp_path, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, (cmd, r.stderr[-1500:])
    return tree.run_tree(n, "srt/batches", outdir, "sahlin", runner=runner, cli=exe, cwd=str(tmp_path))


def _same_from_mindb(a, b):
    """Leaf names (LeftLeaf / RightLeaf) hold the work directory; everything from the MinDB on must be identical."""
    import cerfile
    pa, pb = cerfile.parse(a), cerfile.parse(b)
    keys = ("BatchNr", "BatchStart", "BatchEnd", "BatchBases", "NrCls", "Depth")
    return all(pa[k] == pb[k] for k in keys) and pa["bytes"][pa["mindb_span"][0]:] == pb["bytes"][pb["mindb_span"][0]:]


def test_two_ranks_gloo_tree_equals_serial_tree(tmp_path):
    """world_size 2 over gloo on CPU: ranks split every level, barriers order the levels, the root `.cer` equals the one a
    single process builds.  The jobs run the reference binary (CPU), which also enforces the merge constraints."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built (needs /root/reference)")
    n = _sorted_batches(tmp_path, REF_BIN)
    assert n == 5
    root = _serial_tree(tmp_path, REF_BIN, n, "ser")
    os.makedirs(tmp_path / "par")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, cwd=str(tmp_path), n=n, cli=REF_BIN))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "ROOT par/m_0_4.cer" in r.stdout
    assert _same_from_mindb(str(tmp_path / root.path), str(tmp_path / "par" / "m_0_4.cer"))


@pytest.mark.gpu
def test_tree_through_the_cli_matches_reference_tree(tmp_path):
    """The whole pipeline (sort, 5 leaf jobs, 4 merges) through the B200 CLI == the same tree through the reference."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built")
    (tmp_path / "ref").mkdir()
    (tmp_path / "our").mkdir()
    n = _sorted_batches(tmp_path / "ref", REF_BIN)
    assert _sorted_batches(tmp_path / "our", OUR_BIN) == n == 5
    a = _serial_tree(tmp_path / "ref", REF_BIN, n, "t")
    b = _serial_tree(tmp_path / "our", OUR_BIN, n, "t")
    import cerfile
    assert cerfile.same_file_except_total_reads(str(tmp_path / "ref" / a.path), str(tmp_path / "our" / b.path))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sahlin", "fast"])
def test_tree_subcommand_equals_file_based_tree(tmp_path, mode):
    """`isONclust2 tree` (every batch and every merge in one process, nodes resident on the device, batches rebuilt from the
    decisions) writes the root `.cer` the file-based tree of `cluster` jobs writes — everything but the two leaf names."""
    n = _sorted_batches(tmp_path, OUR_BIN, n_reads=1200, per_batch=200)
    assert n == 6
    from isonclust2_b200 import tree
    os.makedirs(tmp_path / "t", exist_ok=True)

    def runner(cmd, env):
        r = subprocess.run(cmd, cwd=tmp_path, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, (cmd, r.stderr[-1500:])
    root = tree.run_tree(n, "srt/batches", "t", mode, runner=runner, cli=OUR_BIN, cwd=str(tmp_path))
    r = subprocess.run([OUR_BIN, "tree", "-x", mode, "-v", "-o", "root.cer", "srt/batches"], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert _same_from_mindb(str(tmp_path / root.path), str(tmp_path / "root.cer"))
    if os.path.exists(REF_BIN):  # and the reference's own tree over the same sorted batches
        os.makedirs(tmp_path / "r", exist_ok=True)
        rroot = tree.run_tree(n, "srt/batches", "r", mode, runner=runner, cli=REF_BIN, cwd=str(tmp_path))
        assert _same_from_mindb(str(tmp_path / rroot.path), str(tmp_path / "root.cer"))


FAIL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch
import torch.distributed as dist
from isonclust2_b200 import tree
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
def runner(cmd, env):
    out = cmd[cmd.index("-o") + 1]
    if rank == 1 and "c_1.cer" in out:
        raise RuntimeError("job failed on purpose")
    open(os.path.join({cwd!r}, out), "w").write("x")
try:
    tree.run_tree(4, "b", "w", "sahlin", rank, world, runner=runner, cli="none", cwd={cwd!r}, agree=tree.make_agree(dist, torch.device("cpu")))
    sys.stdout.write("RANK%d-FINISHED\n" % rank)
except RuntimeError as e:
    sys.stdout.write("RANK%d-STOPPED: %s\n" % (rank, e))
sys.stdout.flush()
dist.destroy_process_group()
"""


def test_failed_job_stops_every_rank(tmp_path):
    """A job that fails on one rank ends the tree on ALL ranks at the end of that level (no rank is left in a barrier), and the
    failed job leaves no output behind for a later level to pick up."""
    os.makedirs(tmp_path / "w")
    script = tmp_path / "worker.py"
    script.write_text(FAIL_WORKER.format(root=ROOT, cwd=str(tmp_path)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29633", str(script)], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "RANK0-STOPPED" in r.stdout and "RANK1-STOPPED" in r.stdout and "FINISHED" not in r.stdout
    assert os.path.exists(tmp_path / "w" / "c_0.cer") and not os.path.exists(tmp_path / "w" / "c_1.cer")
    assert not [f for f in os.listdir(tmp_path / "w") if f.startswith("m_")]  # no merge level ran


def test_merge_into_purged_left_batch_is_refused(tmp_path):
    """A left batch written with -z has clusters but no MinDB: the reference would not see them, this build would — it refuses
    (before any device work) instead of writing a different merge."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built (needs /root/reference)")
    n = _sorted_batches(tmp_path, REF_BIN, n_reads=400, per_batch=200)
    assert n == 2
    for b, extra in ((0, ["-z"]), (1, [])):
        r = subprocess.run([REF_BIN, "cluster", "-x", "sahlin", *extra, "-l", f"srt/batches/isONbatch_{b}.cer", "-o", f"c_{b}.cer"], cwd=tmp_path,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr[-1500:]
    r = subprocess.run([OUR_BIN, "cluster", "-x", "sahlin", "-l", "c_0.cer", "-r", "c_1.cer", "-o", "m.cer"], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "empty minimizer database" in r.stderr, r.stderr[-1500:]
    assert not os.path.exists(tmp_path / "m.cer")
