"""The C-ABI library loads without a GPU and exports every symbol include/lbm_gpu.h
declares.  No compute call is made here; lattice creation without a device must fail
loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import lbm_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lbm_gpu.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_gpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    assert declared_functions() == sorted(L.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(L.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert L.load_library().lbm_gpu_abi_version() == 2


def test_no_torch_types_or_torch_linkage():
    """plain pointers and sizes only: the library links neither torch nor python."""
    out = subprocess.run(["ldd", L.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "torch" not in out and "python" not in out
    header = open(HEADER).read()
    includes = [l for l in header.splitlines() if l.strip().startswith("#include")]
    assert includes == ["#include <stddef.h>", "#include <stdint.h>"]
    assert "Tensor" not in header and "torch::" not in header and "at::" not in header


def test_param_struct_matches_reference_t_param():
    # t_param: 4 ints + 3 floats, 28 bytes (d2q9-bgk.c:64-73)
    assert C.sizeof(L.Param) == 28
    assert [f[0] for f in L.Param._fields_] == ["nx", "ny", "maxIters", "reynolds_dim", "density", "accel", "omega"]


def test_create_without_gpu_fails_loudly():
    lib = L.load_library()
    if lib.lbm_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(L.LbmError, match="no CUDA device|no CPU fallback"):
        L.Lattice(16, 8, 0.1, 0.005, 1.85)


def test_product_does_not_touch_the_oracle():
    """Nothing under the package, include/ or the Makefile's product targets refers to oracle/."""
    pkg = os.path.join(ROOT, "advanced-hpc-lbm_b200")
    for base, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(base, fn)).read().lower(), fn
    assert "oracle" not in open(HEADER).read().lower()
