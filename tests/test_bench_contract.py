"""bench.py's contract on the CPU: the reference arm (the unmodified reference on the host cores) prints the JSON line the driver
reads, with the keys of the GPU arm's line; the GPU arm refuses to run without a GPU (there is no CPU path to time by accident)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_line():
    from oracle import ref
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-iters", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout                      # exactly ONE line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    if not ref.available():
        assert "unavailable" in d
        return
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "ct_mul/s" and d["unit"] == "ct_mul/s" and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "ct_mul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same workload string as the GPU arm (the driver compares the two configs)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.WORKLOAD
    assert abs(d["ms_per_step"] * d["steps"] * 1e-3 * d["value"] - d["config"]["pairs_per_step"] * d["steps"]) < 1e-6 * d["config"]["pairs_per_step"] * d["steps"] + 1e-6


def test_reference_arm_other_ranks_do_no_work():
    """under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without output"""
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU path" in (r.stdout + r.stderr)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]         # and prints no JSON line
