"""bench.py contract on the CPU: the reference arm runs without a GPU (oracle port on the host cores) and prints the
JSON line the driver parses; the B200 arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "3", "--warmup", "1", "--n-envs", "32")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "humanoid_physics_env_steps_per_sec" and d["unit"] == "physics env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["steps"] == 3 and d["warmup"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a visible GPU")
    r = _run("--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
