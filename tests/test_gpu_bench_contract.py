"""bench.py's GPU arm on a reduced grid (LBM_BENCH_NX / LBM_BENCH_ROWS): the JSON line
carries every key of the benchmark contract and the numbers are self-consistent."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_arm_prints_contract_line():
    env = dict(os.environ, LBM_BENCH_NX="4096", LBM_BENCH_ROWS="4096", LBM_BENCH_CPU_BUDGET_S="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
                        "--timesteps", "20", "--no-cpu-baseline"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["unit"] == "MLUPS" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["scaling"] == "weak" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["kernel_variant"] == 512 and d["gpu_launches"] == 2 * 20 // 2   # two timesteps per kernel launch (K7)
    cells = 4096 * 4096
    assert abs(d["value"] - cells * 20 * 2 / (d["ms_per_step"] * 2 * 1e-3) / 1e6) < 1e-6 * d["value"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["timesteps_per_launch"] == 2 and rf["algorithmic_bytes_per_update"] == 36.0
    assert abs(rf["achieved"] - d["value"] * 36e-3) < 1e-6 * rf["achieved"]
    assert 0.3 < rf["frac"] < 1.3
    assert "traffic_source" in rf and len(rf["library_sha256_16"]) == 16
    par = d["parity"]
    assert par["checked"] and par["ok"] and all(c["strict_lattice_equals_oracle"] for c in par["cases"])
    assert par["mass_conservation_full_grid"]["ok"]
    assert d["strong"]["value"] == d["value"]
    sh = d["shipped"]
    assert set(sh) == {"128x128", "128x256", "256x256", "1024x1024"} and all(v["check"] == "pass" for v in sh.values())
    e = d["e2e"]
    assert e["unit"] == "MLUPS" and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] == cells * 4 and e["d2h_bytes_per_step"] == 4 * cells * 4 + 20 * 4
    assert d["config"]["workload"].startswith("synthetic 4096x4096 channel")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
