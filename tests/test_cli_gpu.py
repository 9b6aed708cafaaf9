"""GPU: the drop-in CLI (isonclust2_b200/bin/isONclust2) against the reference binary (oracle/_ref/isONclust2) on the same
FASTQ with the same arguments: every output file of `sort` and every `.cer` of `cluster` (three modes, -j, -z, no -x, two
merge levels) must be byte-identical.  Both binaries run with the same relative paths from sibling directories, because
paths as typed are part of the `.cer`."""
import filecmp
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "isONclust2")
OUR_BIN = os.path.join(ROOT, "isonclust2_b200", "bin", "isONclust2")

pytestmark = pytest.mark.gpu


def run(exe, args, cwd, ok=True):
    r = subprocess.run([exe] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    if ok:
        assert r.returncode == 0, (exe, args, r.stderr[-2000:])
    return r


def same_cer(a, b):
    """Byte equality; TotalReads (bytes 28..31) of a batch written by `sort` is uninitialised memory in the reference
    (serialize.h:29), inherited by every `cluster` output downstream of it, and is compared masked."""
    import cerfile
    return cerfile.same_file_except_total_reads(a, b)


@pytest.fixture(scope="module")
def both(tmp_path_factory):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built")
    assert os.path.exists(OUR_BIN), "isonclust2_b200/bin/isONclust2 missing: run __graft_entry__.build()"
    from isonclust2_b200 import synth
    base = tmp_path_factory.mktemp("cli")
    dirs = {}
    seq, qual, offs = synth.make_reads(900, seed=3)
    for who in ("ref", "our"):
        d = base / who
        d.mkdir()
        synth.write_fastq(str(d / "reads.fq"), seq, qual, offs)
        dirs[who] = d
    return dirs, {"ref": REF_BIN, "our": OUR_BIN}


def both_run(both, args):
    dirs, exe = both
    for who in ("ref", "our"):
        run(exe[who], args, dirs[who])


def test_sort_outputs_are_byte_identical(both):
    dirs, _ = both
    both_run(both, ["sort", "-B", "50000", "-M", "300", "-o", "srt", "reads.fq"])
    for name in ("sorted_reads.fastq", "sorted_reads_idx.tsv", "sorted_reads_idx.cer", "scores.tsv"):
        assert filecmp.cmp(dirs["ref"] / "srt" / name, dirs["our"] / "srt" / name, shallow=False), name
    names = sorted(os.listdir(dirs["ref"] / "srt" / "batches"))
    assert names == sorted(os.listdir(dirs["our"] / "srt" / "batches")) and len(names) == 3
    for n in names:
        assert same_cer(str(dirs["ref"] / "srt" / "batches" / n), str(dirs["our"] / "srt" / "batches" / n)), n


def _need_sorted(both):
    dirs, _ = both
    if not (dirs["our"] / "srt" / "batches" / "isONbatch_0.cer").exists():
        both_run(both, ["sort", "-B", "50000", "-M", "300", "-o", "srt", "reads.fq"])


@pytest.mark.parametrize("extra,out", [(["-x", "sahlin"], "c0.cer"), (["-x", "fast"], "f0.cer"), (["-x", "furious"], "u0.cer"),
                                       (["-x", "sahlin", "-j"], "j0.cer"), (["-x", "sahlin", "-z"], "z0.cer"), ([], "n0.cer")])
def test_cluster_single_batch_is_byte_identical(both, extra, out):
    dirs, _ = both
    _need_sorted(both)
    both_run(both, ["cluster"] + extra + ["-l", "srt/batches/isONbatch_0.cer", "-o", out])
    assert same_cer(str(dirs["ref"] / out), str(dirs["our"] / out))


def test_merge_tree_is_byte_identical(both):
    """cluster every batch, then merge ((0,1),2): the left MinDB is reloaded, grown and written in container order."""
    dirs, _ = both
    _need_sorted(both)
    for i in range(3):
        both_run(both, ["cluster", "-x", "sahlin", "-l", f"srt/batches/isONbatch_{i}.cer", "-o", f"c{i}.cer"])
        assert same_cer(str(dirs["ref"] / f"c{i}.cer"), str(dirs["our"] / f"c{i}.cer")), i
    both_run(both, ["cluster", "-x", "sahlin", "-l", "c0.cer", "-r", "c1.cer", "-o", "m01.cer"])
    assert same_cer(str(dirs["ref"] / "m01.cer"), str(dirs["our"] / "m01.cer"))
    both_run(both, ["cluster", "-x", "sahlin", "-l", "m01.cer", "-r", "c2.cer", "-o", "m012.cer"])
    assert same_cer(str(dirs["ref"] / "m012.cer"), str(dirs["our"] / "m012.cer"))
    # a chain where the reference's input is OUR output and vice versa gives the same bytes again
    shutil.copy(dirs["ref"] / "m01.cer", dirs["our"] / "m01_ref.cer")
    run(OUR_BIN, ["cluster", "-x", "sahlin", "-l", "m01_ref.cer", "-r", "c2.cer", "-o", "m012b.cer"], dirs["our"])
    import cerfile  # LeftLeaf differs by construction: compare everything from the MinDB on
    pa, pb = cerfile.parse(str(dirs["our"] / "m012b.cer")), cerfile.parse(str(dirs["ref"] / "m012.cer"))
    assert pa["mindb"] == pb["mindb"] and pa["bytes"][pa["mindb_span"][0]:] == pb["bytes"][pb["mindb_span"][0]:]


def test_config0_500_reads_single_batch(tmp_path):
    """BASELINE.json configs[0]: sort + cluster -x sahlin of a 500-read FASTQ, k=11 w=15, one batch."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built")
    from isonclust2_b200 import synth
    seq, qual, offs = synth.make_reads(500, seed=1, reads_per_gene=25)
    for who, exe in (("ref", REF_BIN), ("our", OUR_BIN)):
        d = tmp_path / who
        d.mkdir()
        synth.write_fastq(str(d / "ens500.fq"), seq, qual, offs)
        run(exe, ["sort", "-k", "11", "-w", "15", "-o", "out", "ens500.fq"], d)
        run(exe, ["cluster", "-x", "sahlin", "-l", "out/batches/isONbatch_0.cer", "-o", "clustered.cer"], d)
    assert same_cer(str(tmp_path / "ref/out/batches/isONbatch_0.cer"), str(tmp_path / "our/out/batches/isONbatch_0.cer"))
    assert same_cer(str(tmp_path / "ref/clustered.cer"), str(tmp_path / "our/clustered.cer"))


def test_other_k_w_and_thresholds(tmp_path):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built")
    from isonclust2_b200 import synth
    seq, qual, offs = synth.make_reads(300, seed=7)
    for who, exe in (("ref", REF_BIN), ("our", OUR_BIN)):
        d = tmp_path / who
        d.mkdir()
        synth.write_fastq(str(d / "r.fq"), seq, qual, offs)
        run(exe, ["sort", "-k", "13", "-w", "20", "-q", "8.5", "-m", "4", "-r", "0.7", "-a", "0.3", "-f", "0.75", "-o", "o", "r.fq"], d)
        run(exe, ["cluster", "-x", "sahlin", "-l", "o/batches/isONbatch_0.cer", "-o", "c.cer"], d)
    assert same_cer(str(tmp_path / "ref/o/batches/isONbatch_0.cer"), str(tmp_path / "our/o/batches/isONbatch_0.cer"))
    assert same_cer(str(tmp_path / "ref/c.cer"), str(tmp_path / "our/c.cer"))


def test_error_behaviour_matches(both, tmp_path):
    dirs, exe = both
    for args in (["cluster", "-o", "x.cer"], ["cluster", "-l", "nope.cer", "-o", "x.cer"], ["sort", "-k", "9", "reads.fq"],
                 ["sort", "-x", "bogus", "reads.fq"], ["frobnicate"]):
        rr, ro = run(exe["ref"], args, dirs["ref"], ok=False), run(exe["our"], args, dirs["our"], ok=False)
        assert rr.returncode == ro.returncode == 1, args
