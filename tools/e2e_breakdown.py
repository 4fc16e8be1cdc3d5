#!/usr/bin/env python3
"""Development probe: where does the end-to-end step of bench.py spend its time?
Run alone (1 GPU) or under torch.distributed.run (N GPUs)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbm_b200 as L
from tools.make_inputs import channel_mask
from importlib import import_module

slabs = import_module("advanced-hpc-lbm_b200.slabs")
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NX, ROWS, T = 16384, 16384, int(os.environ.get("T", "200"))
ny = ROWS * world
row0, nrows = L.split_rows(ny, world)[rank]
pin = L.PinnedArray((nrows, NX), np.int32)
pin.array[...] = channel_mask(NX, ny, rows=(row0, row0 + nrows))
out = [L.PinnedArray((nrows, NX), np.float32) for _ in range(4)]
av = np.empty(T, dtype=np.float32)


def bar():
    if dist is not None:
        dist.barrier()


for it in range(5):
    bar()
    t = [time.perf_counter()]
    if world == 1:
        lat = L.Lattice(NX, ny, 0.1, 0.005, 1.85, obstacles=pin.array, flags=L.POOL)
        t.append(time.perf_counter()); t.append(t[-1]); t.append(t[-1])
    else:
        lat = L.Lattice(NX, ny, 0.1, 0.005, 1.85, obstacles=pin.array, slab=(row0, nrows), device_ids=[local],
                        flags=L.POOL)
        t.append(time.perf_counter())
        lat.ipc_connect_all(slabs.gather_descriptors(lat.ipc_export(), world, dist))
        t.append(time.perf_counter())
        bar(); lat.ipc_prepare(); bar()
        t.append(time.perf_counter())
    lat.run(T, out=av)
    t.append(time.perf_counter())
    lat.final_fields(out=[o.array for o in out])
    t.append(time.perf_counter())
    lat.close()
    t.append(time.perf_counter())
    names = ["create", "ipc exch+connect", "prepare+barriers", "run", "final_fields", "destroy"]
    print("rank %d it %d: " % (rank, it) + "  ".join("%s %.1f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])),
          " total %.1f ms" % (1e3 * (t[-1] - t[0])), flush=True)
if dist is not None:
    bar(); dist.destroy_process_group()
