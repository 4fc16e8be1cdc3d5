"""Row-slab decomposition of the lattice over ranks (host-side logic, no GPU needed).

The reference is single-process; its empty "Collate" phase (d2q9-bgk.c:203-213) is
where a multi-rank version would gather.  Here rows are split into contiguous slabs,
one per GPU, neighbours forming a ring because y is periodic (d2q9-bgk.c:2132-2134).
split_rows mirrors split_rows() in csrc/lbm_gpu.cu so that a one-process-per-GPU
launch and the single-process multi-GPU path cut the grid at the same rows.
"""
import numpy as np


def split_rows(ny, n):
    """[(row0, nrows)] for n ranks: ny // n rows each, remainder to the first ranks."""
    if n < 1 or ny < n:
        raise ValueError("need 1 <= n <= ny (ny=%d, n=%d)" % (ny, n))
    base, rem = divmod(ny, n)
    out, r = [], 0
    for i in range(n):
        k = base + (1 if i < rem else 0)
        out.append((r, k))
        r += k
    return out


def ring_neighbours(rank, n):
    """(below, above): the ranks holding row0-1 and row0+nrows, periodic."""
    return (rank - 1) % n, (rank + 1) % n


def accel_row_owner(ny, n):
    """Rank that holds global row ny-2, the row accelerate_flow acts on (d2q9-bgk.c:240)."""
    target = ny - 2
    for i, (r0, k) in enumerate(split_rows(ny, n)):
        if r0 <= target < r0 + k:
            return i
    raise AssertionError


def exchange_descriptors(desc, rank, world, dist=None):
    """All-gather the per-rank IPC descriptors (uint8[IPC_DESC_BYTES]) and return
    (below, above) for this rank.  `dist` is torch.distributed (any backend)."""
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    if world == 1:
        return desc, desc
    import torch
    t = torch.from_numpy(desc.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    below, above = ring_neighbours(rank, world)
    return parts[below].cpu().numpy(), parts[above].cpu().numpy()


def gather_descriptors(desc, world, dist=None):
    """All-gather the per-rank IPC descriptors -> (world, IPC_DESC_BYTES) uint8, for
    lbm_gpu_ipc_connect_all (the library finds its neighbours and the whole-grid facts itself)."""
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    if world == 1:
        return desc.reshape(1, -1)
    import torch
    t = torch.from_numpy(desc.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return np.stack([p.cpu().numpy() for p in parts])


def combine_step_sums(local_sums, local_free, dist=None, world=1):
    """Per-step sums of |u| and free-cell counts added over ranks -> av_vels (float64).

    This is the one reduction of the whole run (the reference's Collate hook): the
    per-step partials stay on each GPU until the run ends."""
    sums = np.asarray(local_sums, dtype=np.float64)
    if world == 1:
        return sums / float(local_free)
    import torch
    t = torch.from_numpy(np.concatenate([sums, [float(local_free)]]))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    t = t.cpu().numpy()
    return t[:-1] / t[-1]
