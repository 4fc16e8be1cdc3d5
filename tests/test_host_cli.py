"""The C host program keeps the reference's command line and error conventions
(d2q9-bgk.c:159-167 argc check / usage, :3001-3013 die / usage, parser messages
:2727-2765 and :2847-2853).  These paths end before any GPU call, so they run on CPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "d2q9-bgk")
PARAMS = os.path.join(ROOT, "inputs", "input_128x128.params")
OBST = os.path.join(ROOT, "inputs", "obstacles_128x128.dat")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="host binary not built (make)")


def run(*args, cwd=None):
    r = subprocess.run([EXE, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=cwd)
    return r.returncode, r.stdout, r.stderr


def test_usage_on_wrong_argument_count():
    for args in ((), (PARAMS,), (PARAMS, OBST, "extra")):
        rc, out, err = run(*args)
        assert rc == 1
        assert err == "Usage: %s <paramfile> <obstaclefile>\n" % EXE


def test_missing_files_die_with_reference_messages(tmp_path):
    rc, _, err = run("/nonexistent.params", OBST)
    assert rc == 1 and err.startswith("Error at line ")
    assert "could not open input parameter file: /nonexistent.params" in err
    rc, _, err = run(PARAMS, "/nonexistent.dat")
    assert rc == 1 and "could not open input obstacles file: /nonexistent.dat" in err


@pytest.mark.parametrize("content,msg", [
    ("", "could not read param file: nx"),
    ("128\n128\nabc\n", "could not read param file: maxIters"),
    ("128\n128\n10\n10\n0.1\n0.005\n", "could not read param file: omega"),
])
def test_bad_param_files(tmp_path, content, msg):
    p = tmp_path / "bad.params"
    p.write_text(content)
    rc, _, err = run(str(p), OBST)
    assert rc == 1 and msg in err and err.startswith("Error at line ")


@pytest.mark.parametrize("content,msg", [
    ("1 2\n", "expected 3 values per line in obstacle file"),
    ("1 2 x\n", "expected 3 values per line in obstacle file"),
    ("128 0 1\n", "obstacle x-coord out of range"),
    ("-1 0 1\n", "obstacle x-coord out of range"),
    ("0 128 1\n", "obstacle y-coord out of range"),
    ("0 0 2\n", "obstacle blocked value should be 1"),
])
def test_bad_obstacle_files(tmp_path, content, msg):
    o = tmp_path / "bad.dat"
    o.write_text("0 0 1\n" + content)
    rc, _, err = run(PARAMS, str(o))
    assert rc == 1 and msg in err


def test_without_gpu_it_dies_instead_of_falling_back(tmp_path):
    import lbm_b200 as L
    if L.load_library().lbm_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    rc, out, err = run(PARAMS, OBST, cwd=str(tmp_path))
    assert rc == 1 and "no CUDA device" in err
    assert not os.path.exists(os.path.join(str(tmp_path), "av_vels.dat"))
