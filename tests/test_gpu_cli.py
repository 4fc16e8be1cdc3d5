"""End to end through the C host program: `./d2q9-bgk <paramfile> <obstaclefile>` on every
shipped input, validated by check/check.py against the golden outputs (expanded from
tests/golden/) at the default 1 % tolerance -- BASELINE.json configs 1-3 -- and, with
LBM_PRECISION=f64, against the golden values to printing precision."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "d2q9-bgk")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import expand_golden  # noqa: E402

NAMES = ["128x128", "128x256", "256x256", "1024x1024"]


def run_cli(name, cwd, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([EXE, os.path.join(ROOT, "inputs", "input_%s.params" % name),
                        os.path.join(ROOT, "inputs", "obstacles_%s.dat" % name)],
                       cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    return r.stdout


def run_check(name, cwd, golden_dir):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "check", "check.py"),
                        "--ref-av-vels-file=" + os.path.join(golden_dir, name + ".av_vels.dat"),
                        "--ref-final-state-file=" + os.path.join(golden_dir, name + ".final_state.dat"),
                        "--av-vels-file=" + os.path.join(cwd, "av_vels.dat"),
                        "--final-state-file=" + os.path.join(cwd, "final_state.dat")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    pcts = [abs(float(x)) for x in re.findall(r"= (\S+)%", r.stdout)]
    return r.returncode, r.stdout, pcts


@pytest.fixture(scope="module")
def golden_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("check"))
    for n in NAMES:
        expand_golden.expand(n, d)
    return d


@pytest.mark.parametrize("name", NAMES)
def test_make_check_passes_fp32(name, tmp_path, golden_dir):
    out = run_cli(name, str(tmp_path))
    lines = out.splitlines()
    assert lines[0] == "==done=="                                   # stdout contract, d2q9-bgk.c:216-221
    assert re.match(r"^Reynolds number:\t\t\d\.\d{12}E[+-]\d\d$", lines[1])
    for i, label in enumerate(["Init", "Compute", "Collate", "Total"]):
        assert re.match(r"^Elapsed %s time:\t\t\t\d+\.\d{6} \(s\)$" % label, lines[2 + i])
    rc, text, pcts = run_check(name, str(tmp_path), golden_dir)
    assert rc == 0 and "Both tests passed!" in text, text
    # fp32 vs the fp64 golden: the reference's own fp32 build lands 0.03-0.14 % away
    # (SURVEY.md section 6); ours must be in the same band
    assert max(pcts) < 0.3, text
    print(name, "av_vels %.3g%% pressure %.3g%%" % tuple(pcts))


@pytest.mark.parametrize("name", NAMES)
def test_f64_kernel_reproduces_golden_files(name, tmp_path, golden_dir):
    """The double-precision build of the kernel is the golden generator's arithmetic: the
    output files agree with check/*.dat to ~1e-10 relative over the whole run (for 256x256
    and 1024x1024 the final_state golden is the regenerated one, tests/golden/make_golden.py)."""
    run_cli(name, str(tmp_path), {"LBM_PRECISION": "f64"})
    rc, text, pcts = run_check(name, str(tmp_path), golden_dir)
    assert rc == 0, text
    assert max(pcts) < 1e-7, text          # per cent, i.e. 1e-9 relative
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    fs = np.loadtxt(os.path.join(str(tmp_path), "final_state.dat"))
    assert np.array_equal(fs[:, 6].reshape(g["obstacle"].shape), g["obstacle"])
    if "u" in g.files:
        assert np.max(np.abs(fs[:, 4].reshape(g["u"].shape) - g["u"])) < 1e-11


def test_multi_slab_cli_gives_identical_files(tmp_path):
    """LBM_GPUS=N must not change a single byte of the outputs (bitwise-identical lattice,
    exact av_vels sums).  Uses as many GPUs as the box has, at least exercising N=1."""
    import lbm_b200 as L
    n = max(1, min(4, L.load_library().lbm_gpu_device_count()))
    a, b = tmp_path / "one", tmp_path / "many"
    a.mkdir(); b.mkdir()
    run_cli("128x256", str(a))
    run_cli("128x256", str(b), {"LBM_GPUS": str(n)})
    for f in ("av_vels.dat", "final_state.dat"):
        assert open(os.path.join(str(a), f), "rb").read() == open(os.path.join(str(b), f), "rb").read()


def test_debug_mode_prints_reference_debug_lines(tmp_path):
    """LBM_DEBUG=1 reproduces the reference's -DDEBUG block (d2q9-bgk.c:196-200) and the total
    density it prints stays constant (mass conservation)."""
    p = tmp_path / "small.params"
    p.write_text("64\n32\n25\n10\n0.1\n0.005\n1.85\n")
    o = tmp_path / "small.dat"
    o.write_text("".join("%d 0 1\n%d 31 1\n" % (x, x) for x in range(64)))
    env = dict(os.environ, LBM_DEBUG="1")
    r = subprocess.run([EXE, str(p), str(o)], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "==timestep: 0==" and lines[3] == "==timestep: 1=="
    dens = [float(l.split(":")[1]) for l in lines if l.startswith("tot density:")]
    avs = [float(l.split(":")[1]) for l in lines if l.startswith("av velocity:")]
    assert len(dens) == 25 and len(avs) == 25
    assert abs(dens[0] - 64 * 32 * 0.1) < 1e-3 and max(dens) - min(dens) < 1e-4
    file_avs = np.loadtxt(os.path.join(str(tmp_path), "av_vels.dat"), usecols=[1])
    np.testing.assert_allclose(avs, file_avs, rtol=1e-12)
    assert "==done==" in lines


def test_json_report_line(tmp_path):
    import json
    out = run_cli("128x128", str(tmp_path), {"LBM_REPORT": "json", "LBM_SKIP_FINAL_STATE": "1"})
    rep = json.loads(out.strip().splitlines()[-1])
    assert rep["nx"] == 128 and rep["ny"] == 128 and rep["steps"] == 40000 and rep["gpus"] == 1
    assert rep["precision"] == "f32" and rep["free_cells"] == 128 * 128 - 508
    assert rep["mlups_device"] > 100 and abs(rep["gbs_72B"] - rep["mlups_device"] * 72e-3) < 0.2
    assert not os.path.exists(os.path.join(str(tmp_path), "final_state.dat"))       # LBM_SKIP_FINAL_STATE
    assert os.path.exists(os.path.join(str(tmp_path), "av_vels.dat"))
