"""bench.py --impl reference: the CPU arm of the benchmark contract runs without a GPU,
prints one JSON line with the keys the driver reads, and times the reference's own
timestep_new2 (oracle/_ref) or, without it, the oracle port."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, LBM_BENCH_CPU_BUDGET_S="1.0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["higher_is_better"] is True
    assert line["value"] > 1.0
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "timesteps" in cb["sample"]
    assert cb["value"] == line["value"]
    assert line["config"]["workload"].startswith("synthetic 16384x16384 channel")
    assert line["gpu_launches"] == 0


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
