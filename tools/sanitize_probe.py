#!/usr/bin/env python3
"""Small run that touches every kernel of liblbm_b200.so; meant to be executed under
`compute-sanitizer --tool memcheck` (one tool per GPU call, B200_PROFILING.md)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lbm_b200 as L
import oracle_lib as O

D, A, W = 0.1, 0.005, 1.85
for (nx, ny) in [(64, 9), (37, 5), (132, 6)]:
    cells, obst = O.random_lattice(nx, ny, seed=1)
    ref, _, _ = O.run(cells, obst, 4, D, A, W)
    kernels = [L.KERNEL_SCALAR, L.KERNEL_PERSISTENT, L.KERNEL_VEC4, L.KERNEL_TMA, L.KERNEL_CLUSTER]
    for k in kernels:
        for n in ((1, 3) if k not in (L.KERNEL_PERSISTENT, L.KERNEL_CLUSTER) else (1,)):
            with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.STRICT | k, n_gpus=n,
                           device_ids=[0] * n) as lat:
                lat.run(3); lat.run(1)
                assert np.array_equal(lat.download(), ref), (nx, ny, k, n)
                lat.final_fields(); lat.av_velocity(); lat.download_rows(1, 2)
                lat.upload(cells); lat.run(1)
    with L.Lattice(nx, ny, D, A, W, obstacles=L.pack_obstacle_bits(obst), bits=True) as lat:
        lat.run(2)
    with L.Lattice(nx, ny, D, A, W, cells=cells.astype(np.float64), obstacles=obst, f64=True) as lat:
        lat.run(2); lat.final_fields()
print("sanitize probe ok")
