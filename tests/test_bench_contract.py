"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU path: the unmodified package from
baseline/_ref or /root/reference when present, else the oracle port) prints ONE JSON line with
the keys the driver reads, only rank 0 prints under a multi-rank launch, and our arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "32", "--cpu-batch", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpoint-iterations/s" and d["unit"] == "Mpoint-iterations/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    sys.path.insert(0, ROOT)
    import bench
    want_kind = "reference" if bench.reference_tree() is not None else "port"     # the unmodified package when it is on the box
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["scaling"] == "strong" and d["config"]["global_batch"] == 256 and d["config"]["n"] == 32
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "32", "--cpu-batch", "1", "--gpus", "2"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_has_no_cpu_path():
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_reference_arm_falls_back_to_the_port(tmp_path):
    """Without the reference package on the box the arm times the oracle port and says so."""
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "32", "--cpu-batch", "1"],
             env={"HELMNET_REFERENCE": str(tmp_path), "HELMNET_BENCH_IGNORE_BASELINE_REF": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_config_is_shared_by_both_arms():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.workload_config(256, 256, 8, "strong")
    assert a["batch_per_gpu"] == 32 and a["global_batch"] == 256 and "C3" in a["workload"] and a["scaling"] == "strong"
