"""ctypes loaders for the CPU checker: our restatement (oracle/liblbm_oracle.so) and,
when it was built in the container that has /root/reference, the compiled reference
itself (oracle/_ref/*.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liblbm_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call([sys.executable, os.path.join(ORACLE_DIR, "build_oracle.py")])
        _oracle = C.CDLL(so)
    return _oracle


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _real(dtype):
    return (C.c_float, "f32") if np.dtype(dtype) == np.float32 else (C.c_double, "f64")


def rest_cells(nx, ny, density, dtype=np.float32):
    ct, sfx = _real(dtype)
    cells = np.empty((ny, nx, 9), dtype=dtype)
    fn = getattr(oracle(), "oracle_init_cells_" + sfx)
    fn.argtypes = [C.c_int, C.c_int, ct, C.c_void_p]
    fn.restype = None
    fn(nx, ny, density, _p(cells))
    return cells


def run(cells, obstacles, iters, density, accel, omega):
    """Oracle step loop.  Returns (final cells, av_vels in REAL, av_vels with double accumulation)."""
    dtype = cells.dtype
    ct, sfx = _real(dtype)
    ny, nx, _ = cells.shape
    a = np.ascontiguousarray(cells).copy()
    b = np.empty_like(a)
    obst = np.ascontiguousarray(obstacles, dtype=np.int32)
    av = np.empty(iters, dtype=dtype)
    avd = np.empty(iters, dtype=np.float64)
    fn = getattr(oracle(), "oracle_run_" + sfx)
    fn.argtypes = [C.c_int, C.c_int, C.c_int, ct, ct, ct, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_void_p]
    fn.restype = C.c_int
    which = fn(nx, ny, iters, density, accel, omega, _p(a), _p(b), _p(obst), _p(av), _p(avd))
    return (a if which == 0 else b), av, avd


def timestep(cells, obstacles, density, accel, omega):
    """One fused oracle step.  Returns (accelerated source, new cells, av in REAL)."""
    dtype = cells.dtype
    ct, sfx = _real(dtype)
    ny, nx, _ = cells.shape
    a = np.ascontiguousarray(cells).copy()
    b = np.empty_like(a)
    obst = np.ascontiguousarray(obstacles, dtype=np.int32)
    fn = getattr(oracle(), "oracle_timestep_" + sfx)
    fn.argtypes = [C.c_int, C.c_int, ct, ct, ct, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = ct
    av = fn(nx, ny, density, accel, omega, _p(a), _p(b), _p(obst), None, None)
    return a, b, av


def timestep_unfused(cells, obstacles, density, accel, omega):
    dtype = cells.dtype
    ct, sfx = _real(dtype)
    ny, nx, _ = cells.shape
    a = np.ascontiguousarray(cells).copy()
    b = np.empty_like(a)
    obst = np.ascontiguousarray(obstacles, dtype=np.int32)
    fn = getattr(oracle(), "oracle_timestep_unfused_" + sfx)
    fn.argtypes = [C.c_int, C.c_int, ct, ct, ct, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = ct
    av = fn(nx, ny, density, accel, omega, _p(a), _p(b), _p(obst))
    return b, av


def final_state(cells, obstacles, density):
    dtype = cells.dtype
    ct, sfx = _real(dtype)
    ny, nx, _ = cells.shape
    obst = np.ascontiguousarray(obstacles, dtype=np.int32)
    out = [np.empty((ny, nx), dtype=dtype) for _ in range(4)]
    fn = getattr(oracle(), "oracle_final_state_" + sfx)
    fn.argtypes = [C.c_int, C.c_int, ct, C.c_void_p, C.c_void_p] + [C.c_void_p] * 4
    fn.restype = None
    fn(nx, ny, density, _p(np.ascontiguousarray(cells)), _p(obst), *[_p(o) for o in out])
    return tuple(out)


def av_velocity(cells, obstacles):
    dtype = cells.dtype
    ct, sfx = _real(dtype)
    ny, nx, _ = cells.shape
    obst = np.ascontiguousarray(obstacles, dtype=np.int32)
    tot = C.c_double()
    n = C.c_long()
    fn = getattr(oracle(), "oracle_av_velocity_" + sfx)
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = ct
    av = fn(nx, ny, _p(np.ascontiguousarray(cells)), _p(obst), C.byref(tot), C.byref(n))
    return av, tot.value, n.value


def calc_reynolds(cells, obstacles, omega, reynolds_dim):
    dtype = cells.dtype
    ct, sfx = _real(dtype)
    ny, nx, _ = cells.shape
    obst = np.ascontiguousarray(obstacles, dtype=np.int32)
    fn = getattr(oracle(), "oracle_calc_reynolds_" + sfx)
    fn.argtypes = [C.c_int, C.c_int, ct, C.c_int, C.c_void_p, C.c_void_p]
    fn.restype = ct
    return fn(nx, ny, omega, reynolds_dim, _p(np.ascontiguousarray(cells)), _p(obst))


# ------------------------------------------------------------------ the real reference
class RefParam(C.Structure):       # t_param, d2q9-bgk.c:64-73
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("maxIters", C.c_int), ("reynolds_dim", C.c_int),
                ("density", C.c_float), ("accel", C.c_float), ("omega", C.c_float)]


class RefParamF64(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("maxIters", C.c_int), ("reynolds_dim", C.c_int),
                ("density", C.c_double), ("accel", C.c_double), ("omega", C.c_double)]


def reference_lib(kind="f32_strict"):
    """The compiled reference (oracle/_ref/libref_<kind>.so) or None if it was not built."""
    so = os.path.join(REF_DIR, "libref_%s.so" % kind)
    if not os.path.exists(so):
        return None
    return C.CDLL(so)


def ref_timestep_new2(lib, cells, obstacles, density, accel, omega, f64=False):
    """Call the reference's own timestep_new2 (d2q9-bgk.c:228).  Returns (src after
    the in-place accelerate, tmp_cells, returned average velocity)."""
    ny, nx, _ = cells.shape
    P = RefParamF64 if f64 else RefParam
    ct = C.c_double if f64 else C.c_float
    p = P(nx, ny, 1, 10, density, accel, omega)
    a = np.ascontiguousarray(cells).copy()
    b = np.zeros_like(a)
    obst = np.ascontiguousarray(obstacles, dtype=np.int32)
    fn = lib.timestep_new2
    fn.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = ct
    av = fn(p, _p(a), _p(b), _p(obst))
    return a, b, av


def ref_call(lib, name, cells, tmp_cells, obstacles, density, accel, omega, f64=False):
    """accelerate_flow / propagate / rebound / collision / av_velocity of the reference."""
    ny, nx, _ = cells.shape
    P = RefParamF64 if f64 else RefParam
    p = P(nx, ny, 1, 10, density, accel, omega)
    fn = getattr(lib, name)
    if name == "accelerate_flow":
        fn.argtypes = [P, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        return fn(p, _p(cells), _p(obstacles))
    if name == "propagate":
        fn.argtypes = [P, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        return fn(p, _p(cells), _p(tmp_cells))
    if name in ("rebound", "collision"):
        fn.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        return fn(p, _p(cells), _p(tmp_cells), _p(obstacles))
    if name == "av_velocity":
        fn.argtypes = [P, C.c_void_p, C.c_void_p]
        fn.restype = C.c_double if f64 else C.c_float
        return fn(p, _p(cells), _p(obstacles))
    raise ValueError(name)


# ------------------------------------------------------------------------ test lattices
def random_lattice(nx, ny, seed, density=0.1, p_obst=0.05, dtype=np.float32, walls=True,
                   obst_on_accel_row=True):
    """A perturbed rest state (all speeds positive) and a random mask with obstacles on
    the wrap edges and (optionally) on row ny-2."""
    rng = np.random.default_rng(seed)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4) * density
    cells = (w[None, None, :] * rng.uniform(0.7, 1.3, size=(ny, nx, 9))).astype(dtype)
    obst = (rng.random((ny, nx)) < p_obst).astype(np.int32)
    if walls and ny > 3:
        obst[0, :] = 1
    if not obst_on_accel_row:
        obst[ny - 2, :] = 0
    obst[:, 0] |= (rng.random(ny) < 0.3).astype(np.int32)
    obst[:, nx - 1] |= (rng.random(ny) < 0.3).astype(np.int32)
    if obst.all():
        obst[ny // 2, nx // 2] = 0
    return cells, obst
