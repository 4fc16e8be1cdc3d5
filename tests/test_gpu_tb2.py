"""K7, the two-timesteps-per-pass kernel (temporal blocking), through the C-ABI.

Two iterations of the reference's loop (d2q9-bgk.c:180-201) fused into one pass over HBM
must give what two separate iterations give:
  * strict build: the oracle's lattice bit for bit and its av_vels, for even and odd step
    counts (an odd last step is done by the one-step kernel), any split of the run into
    calls, several slabs (two-row ghost zones), every segment height;
  * default build: the same BITS as the one-step kernel K1a (both go through the same
    explicit operation sequence), lattice and av_vels.
"""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import lbm_b200 as L
import oracle_lib as O

pytestmark = pytest.mark.gpu
D, A, W = 0.1, 0.005, 1.85


@pytest.fixture
def seg_rows(monkeypatch):
    def set_rows(n):
        if n is None:
            monkeypatch.delenv("LBM_TB2_SEG_ROWS", raising=False)
        else:
            monkeypatch.setenv("LBM_TB2_SEG_ROWS", str(n))
    return set_rows


TWO_STEP = [L.KERNEL_TB2]


@pytest.mark.parametrize("nx,ny", [(512, 8), (520, 11), (1024, 17), (1540, 12), (2048, 70), (516, 130), (1008, 9),
                                   (32, 8), (128, 128), (128, 256), (256, 37), (504, 9), (508, 10), (36, 300)])
@pytest.mark.parametrize("steps", [1, 2, 3, 10])
@pytest.mark.parametrize("kernel", TWO_STEP)
def test_strict_bit_exact(nx, ny, steps, kernel):
    cells, obst = O.random_lattice(nx, ny, seed=nx * 1000 + ny)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.STRICT | kernel) as lat:
        assert lat.info().kernel == kernel
        av = lat.run(steps)
        got = lat.download()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), \
        "%dx%d: %d cells differ" % (nx, ny, np.count_nonzero((got != ref).any(axis=2)))
    np.testing.assert_allclose(av.astype(np.float64), av_ref_d, rtol=1e-7, atol=0)


@pytest.mark.parametrize("rows_per_segment", [2, 3, 5, 7, 64])
def test_every_segment_height(rows_per_segment, seg_rows):
    seg_rows(rows_per_segment)
    nx, ny, steps = 1024, 23, 6
    cells, obst = O.random_lattice(nx, ny, seed=rows_per_segment, p_obst=0.03)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.STRICT | L.KERNEL_TB2) as lat:
        av = lat.run(steps)
        got = lat.download()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    np.testing.assert_allclose(av.astype(np.float64), av_ref_d, rtol=1e-7, atol=0)


@pytest.mark.parametrize("tail_rows", [2, 3, 0])
def test_short_segments_at_the_end_of_a_launch(tail_rows, monkeypatch):
    """Tall slabs end a launch with short segments (so that the SMs drain together): same bits."""
    monkeypatch.setenv("LBM_TB2_TAIL_ROWS", str(tail_rows))
    nx, ny, steps = 1024, 1700, 5
    cells, obst = O.random_lattice(nx, ny, seed=40 + tail_rows, p_obst=0.01)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.STRICT | L.KERNEL_TB2) as lat:
        av = lat.run(steps)
        _, cs = lat.digest()
    assert cs == L.lattice_checksum(ref)
    np.testing.assert_allclose(av.astype(np.float64), av_ref_d, rtol=1e-7, atol=0)


@pytest.mark.parametrize("nx,ny", [(512, 16), (1024, 40), (2052, 33), (128, 128)])
def test_default_build_gives_the_bits_of_the_one_step_kernel(nx, ny):
    steps = 21
    cells, obst = O.random_lattice(nx, ny, seed=3, p_obst=0.02)
    res = []
    for k in (L.KERNEL_VEC4, L.KERNEL_TB2):
        with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=k) as lat:
            av = lat.run(steps)
            res.append((lat.download(), av, lat.final_fields()))
    for other in res[1:]:
        assert np.array_equal(res[0][0].view(np.uint32), other[0].view(np.uint32))
        assert np.array_equal(res[0][1], other[1])
        for a, b in zip(res[0][2], other[2]):
            assert np.array_equal(a, b)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    rel = np.abs(res[1][0].astype(np.float64) - ref) / np.abs(ref)
    assert rel.max() <= 2e-5
    np.testing.assert_allclose(res[1][1].astype(np.float64), av_ref_d, rtol=1e-5, atol=0)


def test_long_run_stays_bit_identical_to_the_one_step_kernel():
    """3001 timesteps of a channel with obstacles: the two-step kernel (several strips, several
    segments, an odd last step) and the one-step kernel end in the same bits."""
    nx, ny, steps = 2052, 300, 3001
    cells, obst = O.random_lattice(nx, ny, seed=9, p_obst=0.01)
    res = []
    for k in (L.KERNEL_VEC4, L.KERNEL_TB2):
        with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=k) as lat:
            av = lat.run(steps)
            res.append((lat.digest(), av))
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1]) and np.all(np.isfinite(res[0][1]))


@pytest.mark.parametrize("kernel", TWO_STEP)
def test_chunked_runs_equal_one_run(kernel):
    """run(a); run(b) == run(a+b) for odd and even pieces: a lone step between two-step passes
    keeps the two-row ghost zones filled."""
    nx, ny = 1024, 24
    cells, obst = O.random_lattice(nx, ny, seed=11)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=kernel) as lat:
        av_all = lat.run(12)
        one = lat.download()
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=kernel) as lat:
        av_parts = np.concatenate([lat.run(1), lat.run(3), lat.run(0), lat.run(5), lat.run(2), lat.run(1)])
        parts = lat.download()
    assert np.array_equal(one, parts)
    assert np.array_equal(av_all, av_parts)


@pytest.mark.parametrize("nx,ny,n", [(512, 24, 3), (1024, 67, 4), (512, 16, 2), (1540, 33, 2)])
def test_slabs_on_one_device_equal_the_oracle(nx, ny, n, seg_rows):
    seg_rows(5)
    steps = 9
    cells, obst = O.random_lattice(nx, ny, seed=nx + ny + n)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, n_gpus=n, device_ids=[0] * n,
                   flags=L.STRICT | L.KERNEL_TB2) as lat:
        assert lat.info().kernel == L.KERNEL_TB2 and lat.info().n_gpus == n
        av = np.concatenate([lat.run(4), lat.run(steps - 4)])
        got = lat.download()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    np.testing.assert_allclose(av.astype(np.float64), av_ref_d, rtol=1e-7, atol=0)


def test_upload_between_runs():
    nx, ny = 512, 20
    cells, obst = O.random_lattice(nx, ny, seed=21)
    other, _ = O.random_lattice(nx, ny, seed=22)
    ref, _, _ = O.run(other, obst, 4, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.STRICT | L.KERNEL_TB2) as lat:
        lat.run(3)
        lat.upload(other)
        lat.run(4)
        assert np.array_equal(lat.download(), ref)


def test_selection_and_refusals():
    with L.Lattice(4096, 4096, D, A, W) as lat:
        assert lat.info().kernel == L.KERNEL_TB2                 # beyond L2, fp32, nx % 4 == 0
    with L.Lattice(4098, 4096, D, A, W) as lat:
        assert lat.info().kernel == L.KERNEL_VEC4                # the bulk copies need nx % 4 == 0
    with L.Lattice(1024, 1024, D, A, W) as lat:
        assert lat.info().kernel == L.KERNEL_PERSISTENT          # lives in L2
    with pytest.raises(L.LbmError, match="two-step kernel"):
        L.Lattice(130, 64, D, A, W, flags=L.KERNEL_TB2)
    with pytest.raises(L.LbmError, match="two-step kernel"):
        L.Lattice(512, 6, D, A, W, flags=L.KERNEL_TB2)
    with pytest.raises(L.LbmError, match="two-step kernel"):
        L.Lattice(512, 20, D, A, W, flags=L.KERNEL_TB2, n_gpus=3, device_ids=[0, 0, 0])   # 6-7 rows per slab


def test_blown_up_lattice_reports_nan():
    nx, ny = 512, 12
    cells, obst = O.random_lattice(nx, ny, seed=2, p_obst=0.0, walls=False)
    cells[4:7, 99:102, :] = 0.0                                  # a hole of zero density: cell (5,100) pulls only zeros, u = 0/0
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.KERNEL_TB2) as lat:
        av = lat.run(4)
    assert np.isnan(av).all(), av


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large],
          derandomize=True)
@given(nxq=st.integers(8, 300), ny=st.integers(8, 40), steps=st.integers(1, 7), seed=st.integers(0, 10 ** 6),
       p_obst=st.sampled_from([0.0, 0.02, 0.2]), slabs=st.integers(1, 3), density=st.sampled_from([0.1, 0.37]),
       accel=st.sampled_from([0.005, 0.05, 1.1]), omega=st.sampled_from([0.7, 1.0, 1.85]), split=st.integers(0, 7),
       walls=st.booleans(), seg=st.sampled_from([2, 3, 4, 9, 64]))
def test_any_configuration_matches_the_oracle(nxq, ny, steps, seed, p_obst, slabs, density, accel, omega, split,
                                              walls, seg):
    nx = 4 * nxq
    n = slabs if ny // slabs >= 8 else 1
    kernel = L.KERNEL_TB2
    os.environ["LBM_TB2_SEG_ROWS"] = str(seg)
    try:
        cells, obst = O.random_lattice(nx, ny, seed=seed, density=density, p_obst=p_obst, walls=walls)
        ref, _, av_ref = O.run(cells, obst, steps, density, accel, omega)
        first = min(split, steps)
        with L.Lattice(nx, ny, density, accel, omega, cells=cells, obstacles=obst, flags=L.STRICT | kernel,
                       n_gpus=n, device_ids=[0] * n) as lat:
            av = np.concatenate([lat.run(first), lat.run(steps - first)])
            got = lat.download()
            _, cs = lat.digest()
    finally:
        del os.environ["LBM_TB2_SEG_ROWS"]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert cs == L.lattice_checksum(ref)
    ok = np.isfinite(av_ref)
    assert np.array_equal(np.isfinite(av), ok)
    np.testing.assert_allclose(av[ok].astype(np.float64), av_ref[ok], rtol=1e-6)
