"""CPU suite: the JSON contract of bench.py's reference arm (the only arm that runs without a GPU) and the helpers the
main arm builds its line from."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "600", "--batch-max-seq", "300",
           "--steps", "1", "--warmup", "0"] + extra
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=e)


def test_reference_arm_line():
    r = _run([])
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "reads/s" and line["higher_is_better"] is True
    assert line["metric"] == "reads/sec clustered (sahlin)" and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "reads" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    # both arms publish what they decided; the reference arm leaves it for the GPU arm's parity_vs_reference
    assert len(line["result_digest"]) == 64 and cb["digest"] == line["result_digest"]
    sys.path.insert(0, ROOT)
    import bench
    ref = json.load(open(bench.DIGEST_FILE))
    assert ref[bench.digest_key(600, 100, "sahlin", 300, 24)]["digest"] == line["result_digest"]


def test_reference_arm_fast_mode_and_sample():
    r = _run(["--mode", "fast", "--sample-reads", "300"])
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["metric"] == "reads/sec clustered (fast)" and line["cpu_baseline"]["sample_reads"] == 300


def test_result_digest_matches_port():
    """The digest both arms print is a function of (clusters per batch, per-slot cluster / strand, ALN_INVOKED): the reference
    arm's digest of a small read set equals the digest of the oracle port's assignment of the same batches."""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench
    import oracle
    from isonclust2_b200 import api, pemp, synth

    if not oracle.have_ref():
        import pytest
        pytest.skip("oracle/_ref not built")
    n, M = 500, 200
    seq, qual, offs = synth.make_reads(n, seed=100, err_set=(0.08,), n_genes=20)
    dt, ncl, kind, dg, nb = bench.reference_step(seq, qual, offs, M, 2, "sahlin")
    P = oracle.port()
    s, e = P.fill_scores(qual, offs, 11)
    o = P.sort_order(s)
    sk = P.sketch_batch(seq, qual, offs, 11, 15)
    status = sk["status"].copy()
    status[api.quality_filtered(e, 7.0).astype(bool)] = 0
    cl, st, ncls, nal = np.full(n, -1, np.int32), np.zeros(n, np.int32), [], 0
    for lo, hi in api.split_batches(o, np.diff(offs), 50000, M):
        pc = P.cluster(seq, offs, sk, s, e, o[lo:hi], 11, 15, 0, pemp.load_p_shared(11, 15), status=status)
        cl[lo:hi], st[lo:hi] = pc["cluster"], pc["strand"]
        ncls.append(pc["n_clusters"])
        nal += pc["n_aligned"]
    assert kind == "reference" and nb == len(ncls) == 3
    assert bench.result_digest(ncls, cl, st, nal) == dg


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_traffic_ratio_and_cpu_share():
    sys.path.insert(0, ROOT)
    import bench

    ratio, src = bench.sketch_traffic_ratio()
    assert src and src.startswith("profiles/") and 0.9 < ratio < 1.5
    os.environ["LOCAL_WORLD_SIZE"] = "4"
    try:
        share, cores = bench.cpu_share()
        assert share == max(2, cores // 4)
    finally:
        del os.environ["LOCAL_WORLD_SIZE"]


def test_reference_arm_tree():
    """--tree: the reference's pipeline with the merge tree over `.cer` files; prints the digest of the root's assignment."""
    r = _run(["--tree", "--batch-max-seq", "150"])
    assert r.returncode == 0, r.stderr[-800:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert "merged as a tree" in line["metric"] and len(line["result_digest"]) == 64 and "4 batches" in line["cpu_baseline"]["sample"]
