"""CPU checks of bench.py's contract pieces that need no GPU: the reference arm (`--impl reference`, the CPU port
of the reference path on the host cores) prints ONE JSON line with the agreed keys, times the SAME basin the GPU
arm runs at N GPUs, and only rank 0 works; the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_line():
    r = run(["--impl", "reference", "--size", "128", "--steps", "3", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "grid-point-updates/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("grid-point-updates/s") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 3 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["global_cells"] == [128, 128] and "workload" in d["config"]


def test_reference_arm_times_the_basin_of_n_gpus_and_only_rank_0_works():
    r = run(["--impl", "reference", "--gpus", "2", "--size", "128", "--steps", "2", "--warmup", "3"],
            env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["config"]["global_cells"] == [128, 256] and d["n_gpus"] == 2      # weak scaling: S x (N S) cells
    assert "128x256" in d["cpu_baseline"]["sample"]
    r = run(["--impl", "reference", "--gpus", "2", "--size", "128", "--steps", "2", "--warmup", "3"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = run(["--size", "64", "--steps", "1"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
