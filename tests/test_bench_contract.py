"""CPU: the measurement contract of bench.py that can be checked without a GPU — the reference arm prints exactly one
JSON line on stdout with the keys the driver reads, and the product arm refuses to run without a device (no CPU
fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env_extra, timeout=600):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=env, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line(tmp_path):
    from oracle import refbin
    if not refbin.available():
        pytest.skip("oracle/_ref reference binary not runnable on this host")
    small = {"FNB_BENCH_N": "20000", "FNB_BENCH_Q": "400", "FNB_DATA_CACHE": str(tmp_path)}
    r = run_bench(["--impl", "reference", "--steps", "2", "--warmup", "1"], small)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "cfg1" in line["config"]["workload"]


def test_reference_arm_under_torchrun_env_only_rank0_prints(tmp_path):
    from oracle import refbin
    if not refbin.available():
        pytest.skip("oracle/_ref reference binary not runnable on this host")
    small = {"FNB_BENCH_N": "20000", "FNB_BENCH_Q": "400", "FNB_DATA_CACHE": str(tmp_path), "RANK": "1", "WORLD_SIZE": "2",
             "LOCAL_RANK": "1"}
    r = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1"], small)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench(["--steps", "1"], {"FNB_BENCH_N": "20000", "FNB_BENCH_Q": "400"}, timeout=300)
    assert r.returncode != 0
    assert r.stdout.strip() == ""
    assert "no CPU path" in r.stderr or "no CUDA device" in r.stderr
