"""check/check.py (written afresh) against the reference's check/check.py on the same
files: same exit status, same report text.  Also the golden expansion round trip."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "check", "check.py")
THEIRS = "/root/reference/check/check.py"
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import expand_golden  # noqa: E402


def _run(script, ref_av, ref_fs, av, fs, extra=()):
    r = subprocess.run([sys.executable, script, "--ref-av-vels-file=" + ref_av, "--ref-final-state-file=" + ref_fs,
                        "--av-vels-file=" + av, "--final-state-file=" + fs, *extra],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.returncode, r.stdout


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("golden"))
    expand_golden.expand("128x128", d)
    ref_av, ref_fs = os.path.join(d, "128x128.av_vels.dat"), os.path.join(d, "128x128.final_state.dat")
    # a perturbed copy: 0.3 % off in av_vels, 2 % off in one pressure
    g = np.load(os.path.join(ROOT, "tests", "golden", "128x128.npz"))
    av = g["av_vels"] * 1.003
    sim_av = os.path.join(d, "sim.av_vels.dat")
    with open(sim_av, "w") as f:
        f.write("".join("%d:\t%.12E\n" % (i, v) for i, v in enumerate(av)))
    lines = open(ref_fs).read().splitlines()
    parts = lines[5000].split()
    parts[5] = "%.12E" % (float(parts[5]) * 1.02)
    bad = list(lines)
    bad[5000] = " ".join(parts)
    sim_fs_bad = os.path.join(d, "sim_bad.final_state.dat")
    open(sim_fs_bad, "w").write("\n".join(bad) + "\n")
    return dict(ref_av=ref_av, ref_fs=ref_fs, sim_av=sim_av, sim_fs_bad=sim_fs_bad, d=d)


def test_pass_and_fail_verdicts(files):
    rc, out = _run(MINE, files["ref_av"], files["ref_fs"], files["sim_av"], files["ref_fs"])
    assert rc == 0 and "Both tests passed!" in out
    rc, out = _run(MINE, files["ref_av"], files["ref_fs"], files["sim_av"], files["sim_fs_bad"])
    assert rc == 1 and "final state failed check" in out and "av_vels failed check" not in out
    rc, out = _run(MINE, files["ref_av"], files["ref_fs"], files["sim_av"], files["ref_fs"], ["--tolerance", "0.1"])
    assert rc == 1 and "av_vels failed check" in out


def test_nan_and_shape_mismatch_fail(files):
    d = files["d"]
    nan_av = os.path.join(d, "nan.av_vels.dat")
    txt = open(files["sim_av"]).read().splitlines()
    txt[7] = "7:\tNAN"
    open(nan_av, "w").write("\n".join(txt) + "\n")
    rc, out = _run(MINE, files["ref_av"], files["ref_fs"], nan_av, files["ref_fs"])
    assert rc == 1
    short = os.path.join(d, "short.av_vels.dat")
    open(short, "w").write("\n".join(txt[:100]) + "\n")
    rc, out = _run(MINE, files["ref_av"], files["ref_fs"], short, files["ref_fs"])
    assert rc == 1 and "Different number of steps" in out


@pytest.mark.skipif(not os.path.exists(THEIRS), reason="reference checker not present")
def test_same_output_as_reference_checker(files):
    for av, fs in ((files["sim_av"], files["ref_fs"]), (files["sim_av"], files["sim_fs_bad"]),
                   (files["ref_av"], files["ref_fs"])):
        mine = _run(MINE, files["ref_av"], files["ref_fs"], av, fs)
        theirs = _run(THEIRS, files["ref_av"], files["ref_fs"], av, fs)
        assert mine[0] == theirs[0]
        strip = lambda s: [l for l in s.splitlines() if "RuntimeWarning" not in l and "diff_pcnt" not in l]
        assert strip(mine[1]) == strip(theirs[1])


@pytest.mark.skipif(not os.path.exists("/root/reference/check/128x256.final_state.dat"), reason="reference goldens not present")
def test_expanded_golden_is_byte_identical_to_reference(tmp_path):
    expand_golden.expand("128x256", str(tmp_path))
    for kind in ("av_vels", "final_state"):
        a = open(os.path.join(str(tmp_path), "128x256.%s.dat" % kind), "rb").read()
        b = open("/root/reference/check/128x256.%s.dat" % kind, "rb").read()
        assert a == b
