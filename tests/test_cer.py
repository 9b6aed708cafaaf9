"""CPU: the `.cer` reader/writer and ClusterSortedReads' bookkeeping of isonclust2_b200/cli (host C++) against files
written by the reference binary (oracle/_ref/isONclust2: sort, cluster in two modes, two merge levels).

The per-read decisions come from the oracle port (initial clustering) or from the reference's own output (merges), so
no device is needed; the GPU tests run the real CLI end to end (tests/test_cli_gpu.py)."""
import os
import subprocess

import numpy as np
import pytest

import cerfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "isONclust2")
K, W = 11, 15


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)
    assert r.returncode == 0, (cmd, r.stderr[-2000:])
    return r


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cer_tool") / "cer_tool")
    run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "isonclust2_b200", "cli"), "-o", exe,
         os.path.join(ROOT, "tests", "cpp", "cer_tool.cpp"), os.path.join(ROOT, "isonclust2_b200", "cli", "cer.cpp")])
    return exe


@pytest.fixture(scope="module")
def refrun(tmp_path_factory):
    """600 synthetic reads -> reference `sort -M 200` (3 batches) -> `cluster` of each -> two merge levels."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built (needs /root/reference)")
    from isonclust2_b200 import synth
    d = tmp_path_factory.mktemp("refrun")
    seq, qual, offs = synth.make_reads(600, seed=1)
    fq = str(d / "reads.fq")
    synth.write_fastq(fq, seq, qual, offs)
    run([REF_BIN, "sort", "-B", "50000", "-M", "200", "-o", "srt", "reads.fq"], cwd=d)
    for i in range(3):
        run([REF_BIN, "cluster", "-x", "sahlin", "-l", f"srt/batches/isONbatch_{i}.cer", "-o", f"c{i}.cer"], cwd=d)
    run([REF_BIN, "cluster", "-x", "fast", "-l", "srt/batches/isONbatch_0.cer", "-o", "f0.cer"], cwd=d)
    run([REF_BIN, "cluster", "-x", "sahlin", "-j", "-l", "srt/batches/isONbatch_1.cer", "-o", "j1.cer"], cwd=d)
    run([REF_BIN, "cluster", "-l", "srt/batches/isONbatch_2.cer", "-o", "n2.cer"], cwd=d)  # no -x: Mode None
    run([REF_BIN, "cluster", "-x", "sahlin", "-l", "c0.cer", "-r", "c1.cer", "-o", "m01.cer"], cwd=d)
    run([REF_BIN, "cluster", "-x", "sahlin", "-l", "m01.cer", "-r", "c2.cer", "-o", "m012.cer"], cwd=d)
    return dict(dir=d, seq=seq, qual=qual, offs=offs)


def test_python_reader_reaches_eof(refrun):
    d = refrun["dir"]
    b = cerfile.parse(str(d / "srt/batches/isONbatch_0.cer"))
    assert b["Depth"] == -1 and b["NrCls"] == 200 == len(b["cls"]) and not b["mindb"] and not b["graphs"]
    c = cerfile.parse(str(d / "c0.cer"))
    assert c["Depth"] == 0 and c["NrCls"] == len(c["cls"]) == len(c["graphs"]) and c["mindb"]
    m = cerfile.parse(str(d / "m012.cer"))
    assert m["Depth"] == 2 and m["BatchEnd"] == 599


def test_roundtrip_sorted_batch_is_byte_identical(refrun, tool):
    d = refrun["dir"]
    for i in range(3):
        src = str(d / f"srt/batches/isONbatch_{i}.cer")
        run([tool, "roundtrip", src, str(d / "rt.cer")])
        assert open(src, "rb").read() == open(d / "rt.cer", "rb").read()


@pytest.mark.parametrize("name", ["c0.cer", "f0.cer", "j1.cer", "n2.cer", "m01.cer", "m012.cer"])
def test_roundtrip_clustered_batch(refrun, tool, name):
    """Clusters, reads, graphs and header byte-identical; the MinDB holds the same postings (its order after a reload is
    that of a rebuilt container, like in the reference, and is pinned by the merge tests below)."""
    d = refrun["dir"]
    run([tool, "roundtrip", str(d / name), str(d / "rt.cer")])
    assert cerfile.same_up_to_mindb_order(cerfile.parse(str(d / name)), cerfile.parse(str(d / "rt.cer")))


def _port_decisions(refrun, port, batch_nr, mode):
    from isonclust2_b200 import pemp
    seq, qual, offs = refrun["seq"], refrun["qual"], refrun["offs"]
    score, err = port.fill_scores(qual, offs, K)
    order = port.sort_order(score)
    sk = port.sketch_batch(seq, qual, offs, K, W)
    sl = order[200 * batch_nr:200 * (batch_nr + 1)]
    res = port.cluster(seq, offs, sk, score, err, sl, K, W, mode, pemp.load_p_shared(K, W))
    dec, ncl = [], 0
    for c, s in zip(res["cluster"], res["strand"]):
        if c < 0:
            dec.append((-2, 0))
        elif c == ncl:
            dec.append((-1, 0))
            ncl += 1
        else:
            dec.append((int(c), int(s)))
    return dec


@pytest.mark.parametrize("batch_nr,mode,name,keep", [(0, 0, "c0.cer", 0), (1, 0, "c1.cer", 0), (0, 1, "f0.cer", 0), (1, 0, "j1.cer", 1)])
def test_initial_clustering_output_is_byte_identical(refrun, tool, port, batch_nr, mode, name, keep):
    """Decisions of the oracle port + the CLI's host bookkeeping and writer == the reference binary's `.cer`."""
    d = refrun["dir"]
    dec = _port_decisions(refrun, port, batch_nr, mode)
    with open(d / "dec.txt", "w") as f:
        f.write("".join(f"{b} {s}\n" for b, s in dec))
    src = f"srt/batches/isONbatch_{batch_nr}.cer"
    run([tool, "apply", src, "-", "dec.txt", "out.cer", str(mode), src, "", str(keep)], cwd=d)
    assert open(d / "out.cer", "rb").read() == open(d / name, "rb").read()


def test_mode_none_every_read_founds_a_cluster(refrun, tool):
    d = refrun["dir"]
    src = "srt/batches/isONbatch_2.cer"
    el = [int(x) for x in run([tool, "eligible", src, "-"], cwd=d).stdout.split()]
    with open(d / "dec.txt", "w") as f:
        f.write("".join("-1 0\n" if e else "-2 0\n" for e in el))
    run([tool, "apply", src, "-", "dec.txt", "out.cer", "3", src, ""], cwd=d)
    assert open(d / "out.cer", "rb").read() == open(d / "n2.cer", "rb").read()


@pytest.mark.parametrize("left,right,out", [("c0.cer", "c1.cer", "m01.cer"), ("m01.cer", "c2.cer", "m012.cer")])
def test_merge_output_is_byte_identical(refrun, tool, left, right, out):
    """Loaded left MinDB (rebuilt by emplace_hint in file order) + appended clusters, moved members, graphs."""
    d = refrun["dir"]
    dec = cerfile.merge_decisions(cerfile.parse(str(d / left)), cerfile.parse(str(d / right)), cerfile.parse(str(d / out)))
    assert any(b >= 0 for b, _ in dec)
    if out == "m01.cer":
        assert any(b == -1 for b, _ in dec)
    with open(d / "dec.txt", "w") as f:
        f.write("".join(f"{b} {s}\n" for b, s in dec))
    run([tool, "apply", left, right, "dec.txt", "out.cer", "0", left, right], cwd=d)
    assert open(d / "out.cer", "rb").read() == open(d / out, "rb").read()


def test_chain_graph_matches_reference_graphs(refrun):
    """Every graph the reference wrote with consensus off is the chain of the cluster representative."""
    c = cerfile.parse(str(refrun["dir"] / "c0.cer"))
    import struct
    for members, g in zip(c["cls"], c["graphs"]):
        n = len(members[0]["raw"]["seq"])
        (num_codes,) = struct.unpack_from("<I", g, 0)
        assert num_codes == len(set(members[0]["raw"]["seq"]))
        (n_nodes,) = struct.unpack_from("<Q", g, 4 + 2 * (8 + 1024))
        assert n_nodes == n


# ---------------------------------------------------------------- `dump` (host only: no device needed)
OUR_BIN = os.path.join(ROOT, "isonclust2_b200", "bin", "isONclust2")


def _tree(d):
    out = {}
    for base, _, files in os.walk(d):
        for f in files:
            p = os.path.join(base, f)
            out[os.path.relpath(p, d)] = open(p, "rb").read()
    return out


@pytest.mark.parametrize("name", ["m012.cer", "c0.cer", "f0.cer", "n2.cer"])
def test_dump_matches_reference(refrun, name):
    """`isONclust2 dump` (mainDump main.cpp:204-236, WriteClusters output.cpp:151-275): every output file — batch_info.tsv,
    clusters_info.tsv, clusters.tsv, cluster_cons.fq, cluster_fastq/<id>.fq — byte-identical to the reference binary's."""
    d = refrun["dir"]
    tag = name.replace(".cer", "")
    run([REF_BIN, "dump", "-i", "srt/sorted_reads_idx.cer", "-o", f"dump_ref_{tag}", name], cwd=d)
    run([OUR_BIN, "dump", "-i", "srt/sorted_reads_idx.cer", "-o", f"dump_our_{tag}", name], cwd=d)
    a, b = _tree(d / f"dump_ref_{tag}"), _tree(d / f"dump_our_{tag}")
    assert sorted(a) == sorted(b) and len(a) > 4
    for k in a:
        assert a[k] == b[k], k


def test_dump_argument_errors(refrun):
    d = refrun["dir"]
    for exe in (REF_BIN, OUR_BIN):
        r = subprocess.run([exe, "dump", "m012.cer"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "Specifying output directory is mandatory!" in r.stderr
        r = subprocess.run([exe, "dump", "-o", "dump_bad", "-i", "nope.cer", "m012.cer"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "Failed to load index nope.cer" in r.stderr


def test_dump_and_info_of_consensus_batches(tmp_path):
    """`.cer` files of a consensus-on run (`sort -c 12 -g 4`: representatives named cons_<batch>_<read>, graphs with many
    sequences, a MinDB with emptied posting lists): `dump` and `info` of this build read them like the reference binary."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/isONclust2 not built (needs /root/reference)")
    import consensus_case as cc
    d = tmp_path
    cc.write_input(str(d))
    cc.run_pipeline(REF_BIN, str(d), steps=cc.STEPS[:5])
    m = cerfile.parse(str(d / "m012.cer"))
    assert any(c[0]["raw"]["name"].startswith(b"cons_") for c in m["cls"])
    for name in ("c0.cer", "m012.cer"):
        tag = name.replace(".cer", "")
        run([REF_BIN, "dump", "-i", "srt/sorted_reads_idx.cer", "-o", f"dump_ref_{tag}", name], cwd=d)
        run([OUR_BIN, "dump", "-i", "srt/sorted_reads_idx.cer", "-o", f"dump_our_{tag}", name], cwd=d)
        a, b = _tree(d / f"dump_ref_{tag}"), _tree(d / f"dump_our_{tag}")
        assert sorted(a) == sorted(b) and len(a) > 4
        for k in a:
            assert a[k] == b[k], (name, k)
        ra = subprocess.run([REF_BIN, "info", name], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        rb = subprocess.run([OUR_BIN, "info", name], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert ra.returncode == rb.returncode == 0 and ra.stdout == rb.stdout and ra.stderr == rb.stderr
