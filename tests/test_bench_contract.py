"""CPU suite: the JSON contract of bench.py's reference arm (the only arm that runs without a GPU) and the helpers the
main arm builds its line from."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "600", "--ref-batch-reads", "300",
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
