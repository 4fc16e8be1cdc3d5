#!/usr/bin/env python3
"""Generate d2q9-bgk input files (param file + ``x y 1`` obstacle file).

File formats are the reference's (d2q9-bgk.c:2736-2762 for the 7-line param file,
:2844-2857 for the obstacle lines).  The four shipped configurations are described
by rule (box walls etc.) and regenerated here instead of being
copied; tests/test_inputs.py checks that the generated masks equal the reference's
files cell for cell when /root/reference is present.

Synthetic large grids (BASELINE.json configs 4-5): a channel -- walls on rows 0 and
ny-1, periodic in x -- with Bernoulli(p) random obstacles from a fixed seed.

Usage:
    tools/make_inputs.py shipped  <outdir>
    tools/make_inputs.py channel  <outdir> --nx 16384 --ny 16384 --iters 200 [--p 0.01]
                                  [--seed 20240229] [--block-accel-row]
"""
import argparse
import os

import numpy as np

SHIPPED = {
    #  name        nx    ny    iters  re  density accel  omega  full rows       full cols
    "128x128":   (128,  128,  40000, 10, 0.1,    0.005, 1.85, (0, 127),       (0, 127)),
    "128x256":   (128,  256,  40000, 10, 0.1,    0.005, 1.85, (127,),         (0, 127)),
    "256x256":   (256,  256,  80000, 10, 0.1,    0.005, 1.85, (0, 255),       (0, 255)),
    "1024x1024": (1024, 1024, 20000, 10, 0.1,    0.01,  1.85, (0, 1023),      (0, 341, 1023)),
}


def shipped_mask(name):
    nx, ny, *_rest, rows, cols = SHIPPED[name]
    m = np.zeros((ny, nx), dtype=np.uint8)
    for r in rows:
        m[r, :] = 1
    for c in cols:
        m[:, c] = 1
    return m


CHUNK_ROWS = 1024


def channel_mask(nx, ny, p=0.01, seed=20240229, block_accel_row=False, rows=None):
    """Walls at rows 0 and ny-1, Bernoulli(p) obstacles elsewhere.

    Row ny-2 (the accelerated row, d2q9-bgk.c:240) is kept free unless
    block_accel_row, in which case it takes the random obstacles like any other row.
    Every block of 1024 rows has its own random stream seeded by (seed, block index), so
    a rank can generate just its rows (`rows=(r0, r1)`) and a grid of the same seed and
    nx but smaller ny equals the larger one apart from its last two rows (used for
    reduced-height CPU replicas of the weak-scaling grids)."""
    r0, r1 = (0, ny) if rows is None else rows
    m = np.zeros((r1 - r0, nx), dtype=np.uint8)
    for c in range(r0 // CHUNK_ROWS, (r1 + CHUNK_ROWS - 1) // CHUNK_ROWS):
        rng = np.random.default_rng([seed, c])
        block = rng.random((CHUNK_ROWS, nx), dtype=np.float32) < p
        a, b = max(r0, c * CHUNK_ROWS), min(r1, (c + 1) * CHUNK_ROWS)
        m[a - r0:b - r0] = block[a - c * CHUNK_ROWS:b - c * CHUNK_ROWS]
    for wall in (0, ny - 1):
        if r0 <= wall < r1:
            m[wall - r0, :] = 1
    if not block_accel_row and r0 <= ny - 2 < r1:
        m[ny - 2 - r0, :] = 0
    return m


def write_params(path, nx, ny, iters, reynolds_dim, density, accel, omega):
    with open(path, "w") as f:
        f.write("%d\n%d\n%d\n%d\n%s\n%s\n%s\n" % (nx, ny, iters, reynolds_dim,
                                               repr(density), repr(accel), repr(omega)))


def write_obstacles(path, mask):
    ys, xs = np.nonzero(mask)
    # one "x y 1" line per blocked cell, row-major
    buf = np.empty((len(xs), 3), dtype=np.int64)
    buf[:, 0] = xs
    buf[:, 1] = ys
    buf[:, 2] = 1
    np.savetxt(path, buf, fmt="%d")


def make_shipped(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, (nx, ny, iters, re, rho, acc, om, _r, _c) in SHIPPED.items():
        write_params(os.path.join(outdir, "input_%s.params" % name), nx, ny, iters, re, rho, acc, om)
        write_obstacles(os.path.join(outdir, "obstacles_%s.dat" % name), shipped_mask(name))


def make_channel(outdir, nx, ny, iters, p, seed, block_accel_row, tag=None):
    os.makedirs(outdir, exist_ok=True)
    tag = tag or "%dx%d" % (nx, ny)
    write_params(os.path.join(outdir, "input_%s.params" % tag), nx, ny, iters, 10, 0.1, 0.005, 1.85)
    write_obstacles(os.path.join(outdir, "obstacles_%s.dat" % tag),
                    channel_mask(nx, ny, p, seed, block_accel_row))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("shipped")
    s.add_argument("outdir")
    c = sub.add_parser("channel")
    c.add_argument("outdir")
    c.add_argument("--nx", type=int, required=True)
    c.add_argument("--ny", type=int, required=True)
    c.add_argument("--iters", type=int, default=200)
    c.add_argument("--p", type=float, default=0.01)
    c.add_argument("--seed", type=int, default=20240229)
    c.add_argument("--block-accel-row", action="store_true")
    c.add_argument("--tag", default=None)
    a = ap.parse_args()
    if a.cmd == "shipped":
        make_shipped(a.outdir)
    else:
        make_channel(a.outdir, a.nx, a.ny, a.iters, a.p, a.seed, a.block_accel_row, a.tag)


if __name__ == "__main__":
    main()
