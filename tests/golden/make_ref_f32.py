#!/usr/bin/env python3
"""Fixtures of the reference's OWN single-precision CPU run on the shipped inputs.

BASELINE.json asks for "the max relative error vs the reference CPU run".  The runs take
up to 13 minutes on one core (1024x1024), so they are made once, in the container that
has /root/reference, with oracle/_ref/d2q9-bgk_ref (the unmodified reference, its own
-Ofast flags, -march=x86-64-v3 -- oracle/build_oracle.py), and stored compactly:

  tests/golden/ref_f32_<name>.npz:  av_vels float32[maxIters], pressure float32[ny,nx],
                                    u float32[ny,nx], reynolds, compute_seconds

Usage: make_ref_f32.py [--run-dir build/ref_f32]   (runs whatever is missing)
"""
import argparse
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
NAMES = {"128x128": (128, 128), "128x256": (128, 256), "256x256": (256, 256), "1024x1024": (1024, 1024)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--run-dir", default=os.path.join(ROOT, "build", "ref_f32"))
    a = ap.parse_args()
    exe = os.path.join(ROOT, "oracle", "_ref", "d2q9-bgk_ref")
    for name, (nx, ny) in NAMES.items():
        d = os.path.join(a.run_dir, name)
        log = os.path.join(d, "run.log")
        if not (os.path.exists(log) and "Elapsed Total" in open(log).read()):
            os.makedirs(d, exist_ok=True)
            with open(log, "w") as f:
                subprocess.check_call([exe, os.path.join(ROOT, "inputs", "input_%s.params" % name),
                                       os.path.join(ROOT, "inputs", "obstacles_%s.dat" % name)], cwd=d, stdout=f)
        text = open(log).read()
        re_num = float(re.search(r"Reynolds number:\s+(\S+)", text).group(1))
        comp = float(re.search(r"Elapsed Compute time:\s+(\S+)", text).group(1))
        av = np.loadtxt(os.path.join(d, "av_vels.dat"), usecols=[1]).astype(np.float32)
        fs = np.loadtxt(os.path.join(d, "final_state.dat"), usecols=[4, 5])
        np.savez_compressed(os.path.join(HERE, "ref_f32_%s.npz" % name), av_vels=av,
                            u=fs[:, 0].reshape(ny, nx).astype(np.float32),
                            pressure=fs[:, 1].reshape(ny, nx).astype(np.float32),
                            reynolds=np.float64(re_num), compute_seconds=np.float64(comp))
        print(name, "Re %.12E compute %.1f s (%.1f MLUPS)" % (re_num, comp, nx * ny * len(av) / comp / 1e6))


if __name__ == "__main__":
    main()
