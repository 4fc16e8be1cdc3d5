"""Row-slab decomposition on the GPU: N slabs must give the SAME BITS as one slab --
lattice, av_vels and final fields -- because every cell sees the same inputs and the
per-step |u| sum is an exact integer accumulation.

  * several slabs on ONE device, ordered by CUDA events: runs on a single-GPU box and
    exercises ghost rows, halo pushes, the owner of the accelerated row ny-2 and uneven
    / one-row slabs;
  * real multi-GPU (needs >= 2 devices): event ordering, the device-side flag protocol,
    and the one-process-per-GPU form over CUDA IPC under torch.distributed.run.
"""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import lbm_b200 as L
import oracle_lib as O

pytestmark = pytest.mark.gpu
D, A, W = 0.1, 0.005, 1.85
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ndev():
    return L.load_library().lbm_gpu_device_count()


def single(nx, ny, cells, obst, steps, flags=0):
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=flags) as lat:
        av = lat.run(steps)
        return lat.download(), av, lat.final_fields()


@pytest.mark.parametrize("nx,ny,n", [(128, 64, 2), (128, 64, 4), (64, 10, 3), (64, 7, 7), (32, 3, 2),
                                     (36, 5, 4), (37, 9, 2), (256, 33, 8)])
def test_slabs_on_one_device_equal_single_slab(nx, ny, n):
    steps = 12
    cells, obst = O.random_lattice(nx, ny, seed=nx + ny + n)
    ref, av_ref, fields_ref = single(nx, ny, cells, obst, steps)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, n_gpus=n, device_ids=[0] * n) as lat:
        av = np.concatenate([lat.run(5), lat.run(steps - 5)])
        got = lat.download()
        fields = lat.final_fields()
        assert lat.info().n_gpus == n
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(av, av_ref)
    for a, b in zip(fields, fields_ref):
        assert np.array_equal(a, b)


def test_slabs_strict_equal_oracle():
    nx, ny, n, steps = 64, 11, 3, 6
    cells, obst = O.random_lattice(nx, ny, seed=5)
    ref, _, _ = O.run(cells, obst, steps, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, n_gpus=n, device_ids=[0] * n,
                   flags=L.STRICT) as lat:
        lat.run(steps)
        assert np.array_equal(lat.download(), ref)


def test_tma_kernel_slabs_equal_single_slab():
    nx, ny, n, steps = 1100, 40, 3, 8
    cells, obst = O.random_lattice(nx, ny, seed=17, p_obst=0.03)
    ref, av_ref, _ = single(nx, ny, cells, obst, steps)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, n_gpus=n, device_ids=[0] * n,
                   flags=L.KERNEL_TMA) as lat:
        av = lat.run(steps)
        assert np.array_equal(lat.download(), ref) and np.array_equal(av, av_ref)


def test_flag_protocol_refuses_a_shared_device():
    with pytest.raises(L.LbmError, match="own GPU"):
        L.Lattice(64, 8, D, A, W, n_gpus=2, device_ids=[0, 0], flags=L.SYNC_FLAGS)


def test_slab_handles_refuse_a_shared_device():
    """Two one-slab handles (the one-process-per-GPU form) on the SAME GPU would spin on each
    other's flags inside kernels that may never be co-scheduled: connect refuses."""
    nx, ny = 64, 8
    a = L.Lattice(nx, ny, D, A, W, slab=(0, 4), device_ids=[0])
    b = L.Lattice(nx, ny, D, A, W, slab=(4, 4), device_ids=[0])
    try:
        with pytest.raises(L.LbmError, match="same physical GPU|must not share"):
            a.ipc_connect(b.ipc_export(), b.ipc_export())
    finally:
        a.close(); b.close()


def test_connect_all_checks_the_descriptor_list():
    """lbm_gpu_ipc_connect_all refuses lists that do not tile the grid, and slabs that share a GPU."""
    nx, ny = 64, 16
    a = L.Lattice(nx, ny, D, A, W, slab=(0, 8), device_ids=[0])
    b = L.Lattice(nx, ny, D, A, W, slab=(8, 8), device_ids=[0])
    c = L.Lattice(nx, ny, D, A, W, slab=(4, 8), device_ids=[0])
    try:
        with pytest.raises(L.LbmError, match="cover the grid"):
            a.ipc_connect_all(np.stack([a.ipc_export()]))
        with pytest.raises(L.LbmError, match="cover the grid"):
            a.ipc_connect_all(np.stack([a.ipc_export(), b.ipc_export(), c.ipc_export()]))
        with pytest.raises(L.LbmError, match="next to this slab"):
            a.ipc_connect_all(np.stack([a.ipc_export(), c.ipc_export()]))       # 16 rows, but nobody holds rows 8..
        with pytest.raises(L.LbmError, match="same physical GPU|must not share"):
            a.ipc_connect_all(np.stack([a.ipc_export(), b.ipc_export()]))
        with pytest.raises(L.LbmError, match="not connected"):
            a.run(1)
    finally:
        a.close(); b.close(); c.close()


@pytest.mark.parametrize("mode", ["events", "flags", "default"])
def test_real_multi_gpu_equals_single(mode):
    n = min(ndev(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nx, ny, steps = 512, 203, 40
    cells, obst = O.random_lattice(nx, ny, seed=1, p_obst=0.02)
    ref, av_ref, _ = single(nx, ny, cells, obst, steps)
    flags = {"flags": L.SYNC_FLAGS, "events": L.SYNC_EVENTS, "default": 0}[mode]
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, n_gpus=n, flags=flags) as lat:
        av = lat.run(steps)
        got = lat.download()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(av, av_ref)


@pytest.mark.parametrize("mode", ["events", "flags"])
def test_real_multi_gpu_two_step_kernel_equals_the_oracle(mode):
    """K7 across GPUs: two-row ghost zones pushed over NVLink, passes ordered by flags / events."""
    n = min(ndev(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nx, ny, steps = 1024, 203, 41
    cells, obst = O.random_lattice(nx, ny, seed=4, p_obst=0.02)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    flags = (L.SYNC_FLAGS if mode == "flags" else L.SYNC_EVENTS) | L.STRICT | L.KERNEL_TB2
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, n_gpus=n, flags=flags) as lat:
        av = np.concatenate([lat.run(7), lat.run(steps - 7)])
        got = lat.download()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    np.testing.assert_allclose(av.astype(np.float64), av_ref_d, rtol=1e-7, atol=0)


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import torch, torch.distributed as dist
    import lbm_b200 as L
    import oracle_lib as O
    from importlib import import_module
    slabs = import_module("advanced-hpc-lbm_b200.slabs")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, steps, mode = %(nx)d, %(ny)d, %(steps)d, %(mode)r
    cells, obst = O.random_lattice(nx, ny, seed=1, p_obst=0.02)
    r0, k = L.split_rows(ny, world)[rank]
    lat = L.Lattice(nx, ny, 0.1, 0.005, 1.85, cells=cells[r0:r0 + k], obstacles=obst[r0:r0 + k],
                    slab=(r0, k), device_ids=[local], flags=%(flags)d)
    if mode == "pair":            # the two-descriptor form: one-step kernel
        below, above = slabs.exchange_descriptors(lat.ipc_export(), rank, world, dist)
        lat.ipc_connect(below, above)
    else:                         # all descriptors: the library may pick the two-step kernel
        lat.ipc_connect_all(slabs.gather_descriptors(lat.ipc_export(), world, dist))
    dist.barrier(); lat.ipc_prepare(); dist.barrier()
    if mode == "short" and rank == 1:
        # this rank is asked for fewer steps: the others must give up, not hang
        lat.run_sums(4)
        open(os.path.join(%(out)r, "rank1_done"), "w").write("ok")
    elif mode == "short":
        import time
        t0 = time.time()
        try:
            lat.run_sums(steps)
            msg = "no error"
        except L.LbmError as e:
            msg = str(e)
        open(os.path.join(%(out)r, "err_%%d.txt" %% rank), "w").write("%%.1f\\n%%s" %% (time.time() - t0, msg))
    else:
        sums = np.concatenate([lat.run_sums(7), lat.run_sums(steps - 7)])
        av = slabs.combine_step_sums(sums, lat.info().local_free_cells, dist, world)
        np.save(os.path.join(%(out)r, "rows_%%d.npy" %% rank), lat.download())
        np.save(os.path.join(%(out)r, "kernel_%%d.npy" %% rank), np.array([lat.info().kernel]))
        if rank == 0:
            np.save(os.path.join(%(out)r, "av.npy"), av)
            if mode == "all":     # connect_all also told the library the whole grid's free cells
                np.save(os.path.join(%(out)r, "av_direct.npy"), lat.run(0))
    if mode == "short":
        dist.barrier()            # rank 1's window must outlive the passes its neighbours still run
    # otherwise no barrier before close, on purpose: lbm_gpu_run returns only when the neighbours are done
    lat.close()
    dist.barrier()
    dist.destroy_process_group()
""")


def _launch(tmp_path, n, nx, ny, steps, mode, flags=0, env=None, timeout=600):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "nx": nx, "ny": ny, "steps": steps, "out": str(tmp_path), "mode": mode,
                                "flags": flags})
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n,
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout,
                          env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("mode,nx,expect", [("pair", 512, "VEC4"), ("all", 512, "TB2"), ("all", 516 + 2, "VEC4")])
def test_one_process_per_gpu_over_ipc_equals_single(tmp_path, mode, nx, expect):
    n = min(ndev(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    ny, steps = 203, 41
    r = _launch(tmp_path, n, nx, ny, steps, mode)
    assert r.returncode == 0, r.stdout[-4000:]
    cells, obst = O.random_lattice(nx, ny, seed=1, p_obst=0.02)
    ref, av_ref, _ = single(nx, ny, cells, obst, steps, flags=L.KERNEL_VEC4)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "rows_%d.npy" % i)) for i in range(n)])
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    av = np.load(os.path.join(str(tmp_path), "av.npy"))
    # the sums are exact integers on the device; only the final double division differs
    np.testing.assert_allclose(av, av_ref.astype(np.float64), rtol=2e-7)
    want = {"VEC4": L.KERNEL_VEC4, "TB2": L.KERNEL_TB2}[expect]
    for i in range(n):
        assert int(np.load(os.path.join(str(tmp_path), "kernel_%d.npy" % i))[0]) == want


def test_a_rank_that_stops_early_makes_the_others_fail_not_hang(tmp_path):
    """The flag protocol has an exit (the reference's convention is die(), d2q9-bgk.c:3001-3007):
    rank 1 runs 4 steps, the others ask for 41 and must come back with an error within the
    time-out instead of spinning inside a kernel for ever."""
    n = min(ndev(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _launch(tmp_path, n, 512, 203, 41, "short", env={"LBM_GPU_SYNC_TIMEOUT_MS": "2000"}, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:]
    assert (tmp_path / "rank1_done").exists()
    for i in range(n):
        if i == 1:
            continue
        secs, msg = (tmp_path / ("err_%d.txt" % i)).read_text().split("\n", 1)
        assert float(secs) < 60.0
        assert "abandoned" in msg, msg
