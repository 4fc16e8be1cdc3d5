"""Host-side logic of the multi-GPU path: row-slab decomposition, ring neighbours and
the once-per-run reduction of the per-step sums, the latter over a real world_size-2
torch.distributed group (gloo, CPU)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import lbm_b200 as L
from importlib import import_module

slabs = import_module("advanced-hpc-lbm_b200.slabs")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ny,n", [(16384, 1), (16384, 8), (131072, 8), (10, 3), (7, 7), (256, 5)])
def test_split_rows_covers_grid_contiguously(ny, n):
    parts = L.split_rows(ny, n)
    assert len(parts) == n
    assert parts[0][0] == 0 and parts[-1][0] + parts[-1][1] == ny
    for (a0, ak), (b0, _) in zip(parts, parts[1:]):
        assert a0 + ak == b0
    sizes = [k for _, k in parts]
    assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def test_split_rows_rejects_more_ranks_than_rows():
    with pytest.raises(ValueError):
        L.split_rows(3, 4)


def test_ring_neighbours_are_periodic():
    assert L.ring_neighbours(0, 4) == (3, 1)
    assert L.ring_neighbours(3, 4) == (2, 0)
    assert L.ring_neighbours(0, 1) == (0, 0)
    assert L.ring_neighbours(1, 2) == (0, 0)


@pytest.mark.parametrize("ny,n", [(16384, 8), (10, 3), (4, 4), (5, 2), (3, 2)])
def test_accel_row_owner_holds_row_ny_minus_2(ny, n):
    owner = slabs.accel_row_owner(ny, n)
    r0, k = L.split_rows(ny, n)[owner]
    assert r0 <= ny - 2 < r0 + k


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, %(root)r)
    import torch.distributed as dist
    from importlib import import_module
    slabs = import_module("advanced-hpc-lbm_b200.slabs")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. descriptor exchange: every rank publishes 256 bytes, gets (below, above)
    desc = np.full(256, rank + 1, dtype=np.uint8)
    below, above = slabs.exchange_descriptors(desc, rank, world, dist)
    b, a = slabs.ring_neighbours(rank, world)
    assert below[0] == b + 1 and above[0] == a + 1 and below.size == 256
    # 2. one reduction of per-step sums at the end of the run
    steps = 5
    local = np.arange(steps, dtype=np.float64) + 10.0 * rank
    free = 100 + rank
    av = slabs.combine_step_sums(local, free, dist, world)
    want = (np.arange(steps) * world + 10.0 * sum(range(world))) / sum(100 + r for r in range(world))
    assert np.allclose(av, want, rtol=0, atol=1e-15), (av, want)
    # 3. all descriptors to every rank, in rank order (what lbm_gpu_ipc_connect_all takes)
    all_descs = slabs.gather_descriptors(desc, world, dist)
    assert all_descs.shape == (world, 256) and [int(d[0]) for d in all_descs] == [r + 1 for r in range(world)]
    # 4. bench.py's reductions over ranks: wrapping 64-bit checksum sum, doubles, all_true
    sys.path.insert(0, %(root)r)
    import bench
    R = bench.Ranks.__new__(bench.Ranks)
    R.rank, R.local_rank, R.world, R.dist = rank, rank, world, dist
    bench._f64 = lambda: __import__("torch").float64
    orig = bench.Ranks._reduce
    def cpu_reduce(self, values, op, dtype):
        import torch
        t = torch.tensor(list(values), dtype=dtype)
        dist.all_reduce(t, op=op)
        return t.tolist()
    bench.Ranks._reduce = cpu_reduce
    x = (0xF123456789ABCDEF + rank * 0x8000000000000001) & (2 ** 64 - 1)
    want64 = sum((0xF123456789ABCDEF + r * 0x8000000000000001) for r in range(world)) & (2 ** 64 - 1)
    assert R.sum_u64(x) == want64
    assert R.sum(1.5 + rank) == sum(1.5 + r for r in range(world)) and R.max(float(rank)) == world - 1
    assert np.array_equal(R.sum_array([1.0, rank]), [world, sum(range(world))])
    assert R.all_true(True) and not R.all_true(rank != 1)
    dist.barrier()
    if rank == 0:
        print("OK")
    dist.destroy_process_group()
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_descriptor_exchange_and_final_reduction_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "OK" in r.stdout
