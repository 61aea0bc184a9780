"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line on stdout with the keys the
driver reads, and the GPU arm's launcher refuses a world-size mismatch."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = {**os.environ, "PYTHONPATH": ROOT}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-600:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    r = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in r, key
    assert r["impl"] == "reference" and r["metric"] == "kv_quant_dequant_hbm_gbs" and r["unit"] == "GB/s"
    assert r["value"] > 0 and r["vs_baseline"] is None and r["higher_is_better"] is True
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"] and "model" not in r["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = {**os.environ, "PYTHONPATH": ROOT, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
