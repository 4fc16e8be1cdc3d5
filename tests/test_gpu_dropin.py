"""The drop-in claim, demonstrated: oracle/_ref/d2q9-bgk_ref_gpu is the REFERENCE's own
program (its main, initialise, calc_reynolds, write_values, die) in which only the step
loop d2q9-bgk.c:180-201 was replaced by the lbm_gpu_* calls of INTEGRATION.md section 1
(oracle/reference_binding.inc, applied with sed at build time, never committed as a
copy).  It must pass the checker, and agree with this repository's own host program."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "d2q9-bgk_ref_gpu")
OURS = os.path.join(ROOT, "d2q9-bgk")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import expand_golden  # noqa: E402


def run(exe, name, cwd):
    r = subprocess.run([exe, os.path.join(ROOT, "inputs", "input_%s.params" % name),
                        os.path.join(ROOT, "inputs", "obstacles_%s.dat" % name)], cwd=cwd,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    return r.stdout


@pytest.mark.skipif(not os.path.exists(REF_GPU), reason="oracle/_ref/d2q9-bgk_ref_gpu not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["128x128", "128x256"])
def test_reference_program_with_the_binding(name, tmp_path):
    a, b, g = tmp_path / "ref_gpu", tmp_path / "ours", tmp_path / "golden"
    for d in (a, b, g):
        d.mkdir()
    out_ref = run(REF_GPU, name, str(a))
    out_ours = run(OURS, name, str(b))
    expand_golden.expand(name, str(g))
    chk = subprocess.run([sys.executable, os.path.join(ROOT, "check", "check.py"),
                          "--ref-av-vels-file=%s/%s.av_vels.dat" % (g, name),
                          "--ref-final-state-file=%s/%s.final_state.dat" % (g, name),
                          "--av-vels-file=%s/av_vels.dat" % a, "--final-state-file=%s/final_state.dat" % a],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert chk.returncode == 0 and "Both tests passed!" in chk.stdout, chk.stdout
    # same library, same kernels -> same av_vels text as our own host program
    assert open(a / "av_vels.dat").read() == open(b / "av_vels.dat").read()
    # the Reynolds line: the reference computes it on the host from the downloaded lattice
    # (serial float sum), we from the device's exact sum -> equal to float rounding
    re_ref = float(re.search(r"Reynolds number:\s+(\S+)", out_ref).group(1))
    re_ours = float(re.search(r"Reynolds number:\s+(\S+)", out_ours).group(1))
    # the reference adds 16 k speeds serially in fp32 (d2q9-bgk.c:2665-2714), the device sum is exact
    assert abs(re_ref - re_ours) <= 1e-5 * abs(re_ours)
    # final_state: the reference's write_values on the downloaded lattice vs our device
    # fields -- same IEEE operations on the same values; only the obstacle column differs
    # where the reference's transposed index does (non-square 128x256)
    fa = np.loadtxt(a / "final_state.dat")
    fb = np.loadtxt(b / "final_state.dat")
    assert np.array_equal(fa[:, :6], fb[:, :6])
    if name == "128x128":
        assert open(a / "final_state.dat").read() == open(b / "final_state.dat").read()
