"""bench.py contract on CPU: the reference arm runs without a GPU and prints ONE JSON line with the agreed keys; the
B200 arm refuses to produce a number without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=dict(os.environ, **(env or {})), timeout=600)


def test_reference_arm_line():
    r = run("--impl", "reference", "--mini", "--steps", "3", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "SpMV GFLOP/s" and d["unit"] == "GFLOP/s"
    assert d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("c5:") and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_env():
    """Ranks other than 0 exit 0 without work; rank 0 ignores the launcher's OMP_NUM_THREADS=1."""
    r = run("--impl", "reference", "--mini", "--steps", "2", "--warmup", "1", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
    r = run("--impl", "reference", "--mini", "--steps", "2", "--warmup", "1", "--gpus", "2",
            env={"RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    d = json.loads(r.stdout.strip())
    assert d["n_gpus"] == 2 and d["config"]["workload"].startswith("c5:") and d["scaling"] == "strong"


def test_b200_arm_needs_a_gpu():
    import singlespmv_b200 as sp
    if sp.device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    r = run("--mini", "--steps", "2", "--no-cpu")
    assert r.returncode != 0 and r.stdout.strip() == ""        # no number without the device
