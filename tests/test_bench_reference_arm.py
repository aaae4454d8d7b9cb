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
