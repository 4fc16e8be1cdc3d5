"""Pins the CPU oracle to the reference ITSELF: oracle/_ref/libref_*.so is the
unmodified d2q9-bgk.c compiled from /root/reference (oracle/build_oracle.py), main
renamed so that its functions can be called.  Bit-for-bit equality on random
lattices, including odd widths, obstacles on the wrap edges and on row ny-2.

Skipped where oracle/_ref was not built (a checkout without /root/reference)."""
import numpy as np
import pytest

import oracle_lib as O

D, A, W = 0.1, 0.005, 1.85
SHAPES = [(4, 2), (5, 3), (16, 8), (37, 11), (64, 64), (128, 33), (3, 7)]


def _lib(kind):
    lib = O.reference_lib(kind)
    if lib is None:
        pytest.skip("oracle/_ref/libref_%s.so not built (needs /root/reference)" % kind)
    return lib


@pytest.mark.parametrize("nx,ny", SHAPES)
def test_timestep_new2_f32_bit_exact(nx, ny):
    lib = _lib("f32_strict")
    cells, obst = O.random_lattice(nx, ny, seed=nx + 100 * ny)
    for _ in range(3):           # a few consecutive steps, feeding the output back
        ra, rb, rav = O.ref_timestep_new2(lib, cells, obst, D, A, W)
        oa, ob, oav = O.timestep(cells, obst, D, A, W)
        assert np.array_equal(ra.view(np.uint32), oa.view(np.uint32))      # accelerated source
        assert np.array_equal(rb.view(np.uint32), ob.view(np.uint32))      # new lattice
        assert np.float32(rav) == np.float32(oav)                          # returned average
        cells = ob


@pytest.mark.parametrize("nx,ny", [(16, 8), (37, 11)])
def test_timestep_new2_f64_bit_exact(nx, ny):
    lib = _lib("f64")
    cells, obst = O.random_lattice(nx, ny, seed=5, dtype=np.float64)
    ra, rb, rav = O.ref_timestep_new2(lib, cells, obst, D, A, W, f64=True)
    oa, ob, oav = O.timestep(cells, obst, D, A, W)
    assert np.array_equal(ra, oa) and np.array_equal(rb, ob) and rav == oav


@pytest.mark.parametrize("nx,ny", [(16, 8), (37, 11), (128, 16)])
def test_semantic_originals_bit_exact(nx, ny):
    """accelerate_flow, propagate, rebound, collision, av_velocity one by one."""
    import ctypes as C
    lib = _lib("f32_strict")
    orc = O.oracle()
    cells, obst = O.random_lattice(nx, ny, seed=77)
    obst = np.ascontiguousarray(obst, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    f = C.c_float

    r_cells, o_cells = cells.copy(), cells.copy()
    O.ref_call(lib, "accelerate_flow", r_cells, None, obst, D, A, W)
    orc.oracle_accelerate_flow_f32.argtypes = [C.c_int, C.c_int, f, f, C.c_void_p, C.c_void_p]
    orc.oracle_accelerate_flow_f32(nx, ny, D, A, p(o_cells), p(obst))
    assert np.array_equal(r_cells, o_cells)
    assert not np.array_equal(r_cells, cells)        # the row really changed

    r_tmp, o_tmp = np.zeros_like(cells), np.zeros_like(cells)
    O.ref_call(lib, "propagate", r_cells, r_tmp, obst, D, A, W)
    orc.oracle_propagate_f32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    orc.oracle_propagate_f32(nx, ny, p(o_cells), p(o_tmp))
    assert np.array_equal(r_tmp, o_tmp)

    O.ref_call(lib, "rebound", r_cells, r_tmp, obst, D, A, W)
    orc.oracle_rebound_f32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    orc.oracle_rebound_f32(nx, ny, p(o_cells), p(o_tmp), p(obst))
    assert np.array_equal(r_tmp, o_tmp) and np.array_equal(r_cells, o_cells)

    O.ref_call(lib, "collision", r_cells, r_tmp, obst, D, A, W)
    orc.oracle_collision_f32.argtypes = [C.c_int, C.c_int, f, C.c_void_p, C.c_void_p, C.c_void_p]
    orc.oracle_collision_f32(nx, ny, W, p(o_cells), p(o_tmp), p(obst))
    assert np.array_equal(r_tmp.view(np.uint32), o_tmp.view(np.uint32))

    rav = O.ref_call(lib, "av_velocity", r_tmp, None, obst, D, A, W)
    oav, _, _ = O.av_velocity(o_tmp, obst)
    assert np.float32(rav) == np.float32(oav)


def test_fast_math_reference_is_close_not_equal():
    """The reference's own Makefile flags (-Ofast) reorder the arithmetic: same
    algorithm, last-bit differences.  States the size of 'not bit-exact' per step."""
    lib = _lib("f32_fast")
    cells, obst = O.random_lattice(64, 32, seed=9)
    _, rb, _ = O.ref_timestep_new2(lib, cells, obst, D, A, W)
    _, ob, _ = O.timestep(cells, obst, D, A, W)
    rel = np.abs(rb.astype(np.float64) - ob) / np.abs(ob)
    assert rel.max() < 5e-6


def test_openmp_annotated_reference_matches_serial_lattice():
    """The CPU baseline timed by bench.py is the reference plus ONE pragma on the row
    loop (d2q9-bgk.c:787).  Under the reference's -Ofast flags gcc generates different
    (reassociated) arithmetic for the outlined loop body, so the two builds agree to the
    last bit or two rather than exactly -- the same size as -Ofast vs -O2."""
    serial, omp = _lib("f32_fast"), _lib("f32_omp")
    cells, obst = O.random_lattice(64, 48, seed=10)
    _, a, av_a = O.ref_timestep_new2(serial, cells, obst, D, A, W)
    _, b, av_b = O.ref_timestep_new2(omp, cells, obst, D, A, W)
    rel = np.abs(a.astype(np.float64) - b) / np.abs(a)
    assert rel.max() < 2e-6
    assert abs(av_a - av_b) < 1e-6 * abs(av_a)


@pytest.mark.parametrize("density,accel,omega", [(0.1, 1.2, 1.85), (0.3, 0.9, 1.0), (0.05, 2.5, 0.6), (1.0, 0.001, 1.99)])
def test_other_parameters_and_the_accelerate_guard(density, accel, omega):
    """Parameters far from the shipped ones.  With a large accel the guard of accelerate_flow
    (f3 - w1 > 0 && f6 - w2 > 0 && f7 - w2 > 0, d2q9-bgk.c:247-249) is false for part of the
    row, which the shipped inputs never exercise."""
    lib = _lib("f32_strict")
    nx, ny = 48, 9
    cells, obst = O.random_lattice(nx, ny, seed=31, density=density)
    w1 = np.float32(density) * np.float32(accel) / np.float32(9)
    row = cells[ny - 2, :, 3]
    if accel in (1.2, 0.9):
        assert ((row - w1) > 0).any() and ((row - w1) <= 0).any()      # both branches taken
    if accel == 2.5:
        assert not ((row - w1) > 0).any()                               # guard false everywhere
    for _ in range(3):
        ra, rb, rav = O.ref_timestep_new2(lib, cells, obst, density, accel, omega)
        oa, ob, oav = O.timestep(cells, obst, density, accel, omega)
        assert np.array_equal(ra.view(np.uint32), oa.view(np.uint32))
        assert np.array_equal(rb.view(np.uint32), ob.view(np.uint32))
        assert np.float32(rav) == np.float32(oav) or (np.isnan(rav) and np.isnan(oav))
        cells = ob
