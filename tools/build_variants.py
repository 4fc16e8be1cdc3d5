#!/usr/bin/env python3
"""Build tuning variants of liblbm_b200.so (same ABI, different compile-time knobs of the
step kernel) into advanced-hpc-lbm_b200/variants/ and, with --run, time each of them on the
16384^2 bench grid (development tool; results go to profiles/r01_kernel_variants.md)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "advanced-hpc-lbm_b200")
OUT = os.path.join(PKG, "variants")
VARIANTS = {
    "base": [],
    "approx1": ["-DLBM_APPROX_MODE=1"],
    "approx2": ["-DLBM_APPROX_MODE=2"],
}


def build():
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for name, defs in VARIANTS.items():
        so = os.path.join(OUT, "liblbm_%s.so" % name)
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
               "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), *defs,
               os.path.join(PKG, "csrc", "lbm_gpu.cu"), "-o", so]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        assert p.wait() == 0, name


def run():
    for name in VARIANTS:
        env = dict(os.environ, LBM_B200_LIB=os.path.join(OUT, "liblbm_%s.so" % name))
        args = sys.argv[sys.argv.index("--run") + 1:] or ["--steps", "40", "--reps", "3", "--kernel", "vec4"]
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py"), *args], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        lines = [l for l in r.stdout.splitlines() if "MLUPS" in l]
        print("%-14s %s" % (name, lines[-1] if lines else r.stdout[-300:]), flush=True)


if __name__ == "__main__":
    if "--run" in sys.argv:
        run()
    else:
        build()
