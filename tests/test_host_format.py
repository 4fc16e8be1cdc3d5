"""advanced-hpc-lbm_b200/host/lbm_io.c: the hand-written "%.12E" conversion used by the
output writers must be byte-identical to printf's on every value (the file contract is
d2q9-bgk.c:2978 and :2993)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "advanced-hpc-lbm_b200", "host")


@pytest.fixture(scope="module")
def io_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("io") / "liblbm_io_test.so")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-fopenmp", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
                           "-I" + HOST, os.path.join(HOST, "lbm_io.c"), "-lm", "-o", so])
    lib = C.CDLL(so)
    lib.lbm_format_e12.argtypes = [C.c_char_p, C.c_double]
    lib.lbm_format_e12.restype = C.c_int
    return lib


def fmt(lib, v):
    buf = C.create_string_buffer(64)
    n = lib.lbm_format_e12(buf, v)
    s = buf.value.decode()
    assert n == len(s)
    return s


def test_special_and_edge_values(io_lib):
    vals = [0.0, -0.0, 1.0, -1.0, 10.0, 9.9999999999995, 9.99999999999949, 0.1, 1e-5, 123456789012.5,
            1234567890123.5, 1e12, 1e13, 9.999999999999e12, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
            float(np.float32(0.1)), float(np.float32(1) / np.float32(3)) * 0.1, 3.333333507180e-02,
            0.5 ** 60, 1 - 2 ** -53, 1.0000000000005, 1.00000000000050004, 2.5e-13 + 1, 1e-32, 1e-40,
            float("inf"), float("-inf")]
    for v in vals:
        assert fmt(io_lib, v) == ("%.12E" % v).replace("INF", "INF"), repr(v)
    assert fmt(io_lib, float("nan")).upper().lstrip("-") == "NAN"


def test_matches_printf_on_random_floats_and_doubles(io_lib):
    rng = np.random.default_rng(12345)
    # the values the program prints: fp32 promoted to double, magnitudes 1e-12 .. 1
    f32 = (10.0 ** rng.uniform(-12, 0.5, 300000) * rng.choice([-1.0, 1.0], 300000)).astype(np.float32)
    f64 = 10.0 ** rng.uniform(-25, 12.9, 200000)
    bits = rng.integers(0, 2 ** 63 - 1, 50000, dtype=np.int64).view(np.float64)      # any finite double
    ties = (rng.integers(10 ** 12, 10 ** 13, 20000).astype(np.float64) + 0.5) / 1e3  # near half-way cases
    for arr in (f32.astype(np.float64), f64, bits[np.isfinite(bits)], ties):
        for v in arr.tolist():
            assert fmt(io_lib, v) == "%.12E" % v, repr(v)


# ---- parsers (initialise's file formats, d2q9-bgk.c:2736-2762 and :2844-2857) -----------
def _bits_to_mask(ptr, nx, ny):
    wpr = (nx + 31) // 32
    words = np.ctypeslib.as_array(ptr, shape=(ny * wpr,)).reshape(ny, wpr)
    return np.unpackbits(words.view(np.uint8).reshape(ny, wpr * 4), axis=1, bitorder="little")[:, :nx]


@pytest.mark.parametrize("text,cells", [
    ("", []),                                                  # no obstacles at all
    ("0 0 1\n3 2 1\n", [(0, 0), (3, 2)]),
    ("0 0 1\n0 0 1\n39 4 1", [(0, 0), (39, 4)]),               # duplicates, no trailing newline
    ("  5 1 1 \r\n\n\n7   3\t1\r\n", [(5, 1), (7, 3)]),         # CRLF, blank lines, tabs
    ("1 1 1 2 2 1 3 3 1", [(1, 1), (2, 2), (3, 3)]),           # fscanf treats any whitespace alike
    ("33 0 1\n32 0 1\n31 0 1\n", [(33, 0), (32, 0), (31, 0)]),  # across a 32-bit word boundary
])
def test_obstacle_parser_accepts_what_fscanf_accepts(io_lib, tmp_path, text, cells):
    nx, ny = 40, 5
    f = tmp_path / "o.dat"
    f.write_text(text)
    io_lib.lbm_read_obstacle_bits.argtypes = [C.c_char_p, C.c_int, C.c_int]
    io_lib.lbm_read_obstacle_bits.restype = C.POINTER(C.c_uint32)
    m = _bits_to_mask(io_lib.lbm_read_obstacle_bits(str(f).encode(), nx, ny), nx, ny)
    want = np.zeros((ny, nx), dtype=np.uint8)
    for x, y in cells:
        want[y, x] = 1
    assert np.array_equal(m, want)


def test_param_parser_reads_float_and_double_views(io_lib, tmp_path):
    import lbm_b200 as L
    f = tmp_path / "p.params"
    f.write_text("128\n256\n40000\n10\n0.1\n0.005\n1.85\n")
    pf, pd = L.Param(), L.ParamF64()
    io_lib.lbm_read_params.argtypes = [C.c_char_p, C.POINTER(L.Param), C.POINTER(L.ParamF64)]
    io_lib.lbm_read_params.restype = None
    io_lib.lbm_read_params(str(f).encode(), C.byref(pf), C.byref(pd))
    assert (pf.nx, pf.ny, pf.maxIters, pf.reynolds_dim) == (128, 256, 40000, 10)
    assert (pd.nx, pd.ny, pd.maxIters, pd.reynolds_dim) == (128, 256, 40000, 10)
    assert np.float32(pf.density) == np.float32(0.1) and pd.density == 0.1      # %f vs %lf
    assert np.float32(pf.accel) == np.float32(0.005) and pd.accel == 0.005
    assert np.float32(pf.omega) == np.float32(1.85) and pd.omega == 1.85


def test_final_state_writer_is_identical_to_fprintf(io_lib, tmp_path):
    """The parallel block writer against Python's own %-formatting of the same values (the
    reference's fprintf line, d2q9-bgk.c:2978), on a block with more rows than pieces."""
    nx, nrows, row0 = 37, 150, 1000
    rng = np.random.default_rng(5)
    f = [(10.0 ** rng.uniform(-9, 0, nrows * nx) * rng.choice([-1.0, 1.0], nrows * nx)).astype(np.float32).astype(np.float64)
         for _ in range(4)]
    ny = row0 + nrows
    mask = (rng.random((ny, nx)) < 0.2)
    bits = np.zeros((ny, (nx + 31) // 32), dtype=np.uint32)
    ys, xs = np.nonzero(mask)
    np.bitwise_or.at(bits, (ys, xs // 32), (np.uint32(1) << (xs % 32).astype(np.uint32)))
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = str(tmp_path / "fs.dat")
    fp = libc.fopen(path.encode(), b"w")
    io_lib.lbm_write_final_state_rows.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong] + [C.c_void_p] * 5
    io_lib.lbm_write_final_state_rows.restype = None
    io_lib.lbm_write_final_state_rows(fp, nx, row0, nrows, *[a.ctypes.data_as(C.c_void_p) for a in f],
                                      bits.ctypes.data_as(C.c_void_p))
    libc.fclose(fp)
    want = "".join("%d %d %.12E %.12E %.12E %.12E %d\n" % (i, row0 + r, f[0][r * nx + i], f[1][r * nx + i], f[2][r * nx + i],
                                                        f[3][r * nx + i], int(mask[row0 + r, i]))
                   for r in range(nrows) for i in range(nx))
    assert open(path).read() == want
