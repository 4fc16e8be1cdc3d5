#!/usr/bin/env python3
"""SURVEY.md section 8d, weak-scaling validation: the largest grid ONE GPU can hold
(16384 x 65536 = 1.07 G cells, 77 GB of lattice) run on 1 GPU and on N GPUs must end in
the same lattice, bit for bit (compared through the exact device checksum), with the
same av_vels bits and conserved mass.  Run on a box with N >= 2 GPUs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbm_b200 as L
from tools.make_inputs import channel_mask

nx, ny, steps = 16384, int(os.environ.get("NY", "65536")), int(os.environ.get("STEPS", "12"))
ndev = L.load_library().lbm_gpu_device_count()
t0 = time.time()
bits = L.pack_obstacle_bits(channel_mask(nx, ny))
print("mask %.1f s, %d GPUs visible" % (time.time() - t0, ndev), flush=True)
results = {}
# 1 GPU with the default kernel (K7, two timesteps per pass) and with the one-step kernel K1a; N GPUs
# ordered by CUDA events and by the device-side flag protocol (the default)
configs = [(1, 0, "1 GPU"), (1, L.KERNEL_VEC4, "1 GPU, one-step kernel")]
if ndev > 1:
    configs += [(ndev, L.SYNC_EVENTS, "%d GPUs, events" % ndev), (ndev, 0, "%d GPUs, flags" % ndev)]
for n, flags, label in configs:
    with L.Lattice(nx, ny, 0.1, 0.005, 1.85, obstacles=bits, bits=True, n_gpus=n, flags=flags) as lat:
        m0, _ = lat.digest()
        av = np.concatenate([lat.run(5), lat.run(steps - 5)])
        ms = lat.info().last_run_device_ms / (steps - 5)
        m1, cs = lat.digest()
        kern = lat.info().kernel
    results[label] = (cs, av.copy(), m0, m1)
    print("%-24s kernel %4d  checksum %016x  mass %.6f -> %.6f  %.3f ms/step  %.0f MLUPS" %
          (label, kern, cs, m0, m1, ms, nx * ny / ms / 1e3), flush=True)
ref = results["1 GPU"]
ok = True
for label, (cs, av, m0, m1) in results.items():
    same = cs == ref[0] and np.array_equal(av.view(np.uint32), ref[1].view(np.uint32)) and m1 == ref[3]
    ok &= same and abs(m1 - m0) / m0 < 1e-6
    print("%-24s %s" % (label, "identical to 1 GPU" if same else "DIFFERS"))
print("PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
