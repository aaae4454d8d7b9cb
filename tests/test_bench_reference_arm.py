"""`bench.py --impl reference` on the CPU: the arm must run without the CUDA library, at the size it is asked for, and
print the contract's JSON line (impl, metric, unit, cpu_baseline, e2e with zero copy bytes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_runs_without_the_cuda_library():
    env = dict(os.environ)
    env["MSM_B200_LIB"] = "/nonexistent/libmsm_b200.so"  # loading the product library would fail loudly
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2n", "11", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "bls12_377_g1_msm_throughput" and line["unit"] == "Mpoints/s"
    assert line["config"]["points_per_step"] == 1 << 11 and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(line["value"] - (1 << 11) / (line["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * line["value"]
    assert "reference_table" in line["windows"] and "best" in line["windows"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--log2n", "10",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_our_arm_prints_the_contract_line():
    """A small run of the CUDA arm: every key the contract names is there and the numbers are consistent."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--log2n", "16", "--steps", "4", "--warmup", "3",
                        "--no-strong"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["ok"] is True and d["n_gpus"] == 1 and d["steps"] == 4 and d["gpu_launches"] > 0
    assert abs(d["value"] - (1 << 16) / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 16) * 32 and d["e2e"]["value"] < d["value"] * 1.05
    rf = d["roofline"]
    assert rf["bound"] == "imad" and 0 < rf["frac"] < 1.2 and rf["peak"] > 5 and rf["pure_imad_peak"] > rf["peak"]
    assert d["cpu_baseline"]["kind"] == "port" and "identical: True" in d["cpu_baseline"]["sample"]
    assert d["pipelined"]["same_points_as_sequential"] and d["pipelined"]["c_api_same_points"]
