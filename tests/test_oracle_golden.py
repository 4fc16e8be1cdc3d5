"""Pins the CPU oracle to the reference's golden vectors (tests/golden/*.npz, made from
check/*.dat of the reference by tests/golden/make_golden.py).

The golden files are a DOUBLE precision run of the reference algorithm (SURVEY.md
section 0 fact 9), printed with 13 significant digits; the f64 instantiation of the
oracle must reproduce them to that printing precision (5e-13 relative)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from tools.make_inputs import SHIPPED, shipped_mask

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRINT_EPS = 6e-13       # half a unit of the 13th significant digit, relative


def _setup(name, dtype):
    nx, ny, iters, re, rho, acc, om, _r, _c = SHIPPED[name]
    obst = shipped_mask(name).astype(np.int32)
    cells = O.rest_cells(nx, ny, rho, dtype)
    return nx, ny, iters, rho, acc, om, obst, cells


@pytest.mark.parametrize("name,steps", [("128x256", 3000), ("256x256", 2000), ("1024x1024", 150)])
def test_f64_oracle_reproduces_golden_av_vels_prefix(name, steps):
    nx, ny, iters, rho, acc, om, obst, cells = _setup(name, np.float64)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    _, av, _ = O.run(cells, obst, steps, rho, acc, om)
    rel = np.abs(av - g["av_vels"][:steps]) / g["av_vels"][:steps]
    assert rel.max() <= PRINT_EPS


@pytest.mark.parametrize("name", ["128x128", "128x256"])
def test_f64_oracle_reproduces_golden_full_run(name):
    """Whole 40000-step runs of the two inputs whose golden final_state the reference ships:
    every av_vels value and all four final_state fields (128x256 exercises the y-wrap:
    rows 0 and 255 are open)."""
    nx, ny, iters, rho, acc, om, obst, cells = _setup(name, np.float64)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fin, av, _ = O.run(cells, obst, iters, rho, acc, om)
    assert np.max(np.abs(av - g["av_vels"]) / g["av_vels"]) <= PRINT_EPS
    ux, uy, u, p = O.final_state(fin, obst, rho)
    assert np.max(np.abs(p - g["pressure"]) / g["pressure"]) <= PRINT_EPS
    # velocities: 13 significant digits of values ~1e-2..1e-9 -> absolute
    for mine, key in ((ux, "u_x"), (uy, "u_y"), (u, "u")):
        assert np.max(np.abs(mine - g[key])) <= 1e-14 + PRINT_EPS * np.max(np.abs(g[key]))
    # last column of the golden final_state is the cell's own obstacle flag
    assert np.array_equal(g["obstacle"], obst.astype(np.uint8))


def test_f32_oracle_is_within_check_tolerance_of_golden():
    """The shipped precision: fp32 lands 0.03-0.15 % from the fp64 golden,
    far inside check.py's 1 %.  2000-step prefix of 128x128."""
    name = "128x128"
    nx, ny, iters, rho, acc, om, obst, cells = _setup(name, np.float32)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    _, av, avd = O.run(cells, obst, 2000, np.float32(rho), np.float32(acc), np.float32(om))
    rel = np.abs(avd - g["av_vels"][:2000]) / g["av_vels"][:2000]
    assert rel.max() < 2e-3


def test_fused_and_unfused_steps_agree():
    """timestep_new2 (fused) == accelerate; propagate; rebound; collision; av_velocity
    (the dead timestep_old order, d2q9-bgk.c:1824-1831): same lattice bit for bit."""
    cells, obst = O.random_lattice(40, 13, seed=21)
    _, fused, av1 = O.timestep(cells, obst, 0.1, 0.005, 1.85)
    unfused, av2 = O.timestep_unfused(cells, obst, 0.1, 0.005, 1.85)
    assert np.array_equal(fused, unfused)
    assert av1 == av2


def test_golden_masks_match_generated_inputs():
    for name in SHIPPED:
        g = np.load(os.path.join(GOLD, name + ".npz"))
        assert np.array_equal(g["obstacle"], shipped_mask(name)), name
