#!/usr/bin/env python3
"""Rewrite the golden text files (<name>.av_vels.dat, <name>.final_state.dat) from the
compact fixtures in tests/golden/*.npz, in the reference's formats
(d2q9-bgk.c:2978 and :2993), for check/check.py and `make check`.

Usage: expand_golden.py <outdir> [name ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["128x128", "128x256", "256x256", "1024x1024"]


def fmt_e(v):
    """C's %.12E for an array of doubles -> array of str."""
    return np.char.upper(np.char.mod("%.12e", v))


def expand(name, outdir):
    z = np.load(os.path.join(HERE, name + ".npz"))
    os.makedirs(outdir, exist_ok=True)
    av = z["av_vels"]
    with open(os.path.join(outdir, name + ".av_vels.dat"), "w") as f:
        f.write("".join("%d:\t%s\n" % (i, s) for i, s in enumerate(fmt_e(av))))
    p = z["pressure"]
    ny, nx = p.shape
    zero = np.zeros_like(p)
    ux = z["u_x"] if "u_x" in z.files else zero
    uy = z["u_y"] if "u_y" in z.files else zero
    u = z["u"] if "u" in z.files else zero
    ob = z["obstacle"]
    cols = [fmt_e(a.ravel()) for a in (ux, uy, u, p)]
    ii = np.tile(np.arange(nx), ny)
    jj = np.repeat(np.arange(ny), nx)
    with open(os.path.join(outdir, name + ".final_state.dat"), "w") as f:
        obr = ob.ravel()
        lines = ["%d %d %s %s %s %s %d\n" % (ii[n], jj[n], cols[0][n], cols[1][n], cols[2][n], cols[3][n], obr[n])
                 for n in range(nx * ny)]
        f.write("".join(lines))


def main():
    outdir = sys.argv[1]
    for name in (sys.argv[2:] or NAMES):
        expand(name, outdir)


if __name__ == "__main__":
    main()
