"""bench.py contract pieces that can be checked without a GPU: the reference arm (the CPU restatement timed on the host
cores) prints one JSON line with the agreed keys, and the product arm refuses to run -- loudly -- when there is no B200
(there is no CPU fallback to fall back to)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["metric"] == "diagonal-band pixels scored/sec" and b["unit"] == "pixels/s"
    assert b["value"] > 0 and b["higher_is_better"] is True and b["n_gpus"] == 1 and b["steps"] == 1
    assert b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert b["cpu_baseline"]["value"] == b["value"] and b["cpu_baseline"]["sample"]
    assert b["e2e"] == {"value": b["value"], "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in b["config"] and b["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    from hicpeaks_b200 import _capi
    if _capi.device_count() > 0:
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline")
    assert r.returncode != 0 and "EngineError" in r.stderr and r.stdout.strip() == ""
