"""GPU: clustering with consensus (`sort -c <positive>`) through the drop-in CLI: every `.cer` — leaf batches,
two merge levels, the three spoa alignment types — must be byte-identical to what the REFERENCE binary writes.

  golden   tests/golden/consensus_v1.json: digests of the reference's files (tools/make_consensus_golden.py)
  live     with ICB_CONS_LIVE=1 and oracle/_ref present, the reference binary is run beside ours and the files compared

What runs where: partial-order alignment and consensus on the host (icb_poa_*), the consensus' sketch, its place in the read
set, the index update and every later query / alignment against it on the device."""
import json
import os

import pytest

import consensus_case as cc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "isONclust2")
OUR_BIN = os.path.join(ROOT, "isonclust2_b200", "bin", "isONclust2")

pytestmark = pytest.mark.gpu


def test_consensus_pipeline_matches_reference(tmp_path):
    with open(os.path.join(ROOT, "tests", "golden", "consensus_v1.json")) as f:
        gold = json.load(f)
    assert gold["sort_args"] == cc.SORT_ARGS and gold["reads"] == [cc.N_READS, cc.SEED, cc.READS_PER_GENE]
    d = tmp_path / "our"
    d.mkdir()
    cc.write_input(str(d))
    ours = cc.run_pipeline(OUR_BIN, str(d))
    bad = [k for k in gold["digests"] if ours.get(k) != gold["digests"][k]]
    assert not bad, bad
    if os.environ.get("ICB_CONS_LIVE") == "1" and os.path.exists(REF_BIN):
        import cerfile
        r = tmp_path / "ref"
        r.mkdir()
        cc.write_input(str(r))
        cc.run_pipeline(REF_BIN, str(r))
        for name, _ in cc.STEPS:
            assert cerfile.same_file_except_total_reads(str(d / name), str(r / name)), name


def test_consensus_graph_life_cycle(tmp_path):
    """ConsMinSize 15 / ConsMaxSize 40 / ConsPeriod 60 on four clusters of 75 reads: several reads per device call while no graph
    is near ConsMinSize, ConsPurge restarts, clusters that stop updating for good (cluster.cpp:267-271)"""
    with open(os.path.join(ROOT, "tests", "golden", "consensus_v1.json")) as f:
        gold = json.load(f)["period_case"]
    assert gold["sort_args"] == cc.P_SORT_ARGS and gold["reads"] == cc.P_READS
    cc.write_input(str(tmp_path), dict(cc.P_READS))
    ours = cc.run_pipeline(OUR_BIN, str(tmp_path), steps=cc.P_STEPS, sort_args=cc.P_SORT_ARGS, n_batches=1)
    assert ours == gold["digests"]


def test_consensus_counts_are_reported(tmp_path):
    """-v prints the reference's counters (main.cpp:355-359); consensus must have been invoked and the spoa line names the algorithm"""
    import subprocess
    d = tmp_path
    cc.write_input(str(d))
    subprocess.run([OUR_BIN] + cc.SORT_ARGS, cwd=d, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    r = subprocess.run([OUR_BIN, "cluster", "-v", "-x", "sahlin", "-l", "srt/batches/isONbatch_0.cer", "-o", "v0.cer"], cwd=d,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Generating consensus using spoa algorithm: semi-global" in r.stderr
    line = [ln for ln in r.stderr.splitlines() if ln.startswith("Consensus invocation count: ")]
    assert line and int(line[0].split()[3]) > 0, r.stderr[-1000:]
    with open(os.path.join(ROOT, "tests", "golden", "consensus_v1.json")) as f:
        gold = json.load(f)
    assert cc.digest(str(d / "v0.cer")) == gold["digests"]["c0.cer"]
