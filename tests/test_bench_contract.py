"""bench.py's reference arm (the CPU twin of the path on host cores) runs without a GPU: the JSON line carries the contract's
keys, and under torchrun only rank 0 works and prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_line():
    res = _run({})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"].startswith("rendered-images/sec") and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"] and "train_magicpony_horse" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "16 images" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    res = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm fails loudly instead of timing something else."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu"],
                         capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in res.stderr and res.stdout.strip() == ""
