"""K9, the small-grid persistent kernel that meets at a grid barrier once per TWO timesteps
(the reference's loop d2q9-bgk.c:180-201 on its own shipped grid sizes), through the C-ABI:
strict build bit-exact against the oracle for even and odd step counts, every tile height,
any split of the run into calls; default build bit-identical to the one-step kernel."""
import numpy as np
import pytest

import lbm_b200 as L
import oracle_lib as O

pytestmark = pytest.mark.gpu
D, A, W = 0.1, 0.005, 1.85


@pytest.mark.parametrize("nx,ny", [(128, 128), (128, 256), (256, 256), (32, 8), (64, 7), (256, 37), (4, 6), (8, 4),
                                   (36, 300), (252, 150), (128, 700)])
@pytest.mark.parametrize("steps", [1, 2, 3, 10])
def test_strict_bit_exact(nx, ny, steps):
    cells, obst = O.random_lattice(nx, ny, seed=nx * 1000 + ny)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.STRICT | L.KERNEL_PAIRS) as lat:
        assert lat.info().kernel == L.KERNEL_PAIRS
        av = lat.run(steps)
        got = lat.download()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), \
        "%dx%d: %d cells differ" % (nx, ny, np.count_nonzero((got != ref).any(axis=2)))
    np.testing.assert_allclose(av.astype(np.float64), av_ref_d, rtol=1e-7, atol=0)


@pytest.mark.parametrize("tile_rows", [1, 2, 3, 4])
def test_every_tile_height(tile_rows, monkeypatch):
    monkeypatch.setenv("LBM_PAIRS_TILE_ROWS", str(tile_rows))
    nx, ny, steps = 128, 131, 9
    cells, obst = O.random_lattice(nx, ny, seed=tile_rows, p_obst=0.03)
    ref, _, av_ref_d = O.run(cells, obst, steps, D, A, W)
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.STRICT | L.KERNEL_PAIRS) as lat:
        av = np.concatenate([lat.run(4), lat.run(1), lat.run(0), lat.run(steps - 5)])
        got = lat.download()
        fields = lat.final_fields()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    np.testing.assert_allclose(av.astype(np.float64), av_ref_d, rtol=1e-7, atol=0)
    for a, b in zip(fields, O.final_state(ref, obst, D)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("nx,ny", [(128, 128), (256, 64)])
def test_default_build_gives_the_bits_of_the_other_kernels(nx, ny):
    steps = 33
    cells, obst = O.random_lattice(nx, ny, seed=5, p_obst=0.02)
    res = []
    for k in (L.KERNEL_VEC4, L.KERNEL_PAIRS, L.KERNEL_PERSISTENT):
        with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=k) as lat:
            av = lat.run(steps)
            res.append((None, av, lat.download()))
    for other in res[1:]:
        assert np.array_equal(res[0][2].view(np.uint32), other[2].view(np.uint32))
        assert np.array_equal(res[0][1], other[1])


def test_selection_and_refusals():
    with L.Lattice(128, 128, D, A, W) as lat:
        assert lat.info().kernel == L.KERNEL_PAIRS                 # the reference's smallest shipped grid
    with L.Lattice(256, 256, D, A, W) as lat:
        assert lat.info().kernel == L.KERNEL_PAIRS
    with L.Lattice(130, 64, D, A, W) as lat:
        assert lat.info().kernel == L.KERNEL_PERSISTENT            # ragged width: one barrier per timestep
    with L.Lattice(512, 512, D, A, W) as lat:
        assert lat.info().kernel == L.KERNEL_PERSISTENT            # wider than a block's rows
    with L.Lattice(128, 128, D, A, W, f64=True) as lat:
        assert lat.info().kernel == L.KERNEL_PERSISTENT
    with pytest.raises(L.LbmError, match="pairs kernel"):
        L.Lattice(512, 64, D, A, W, flags=L.KERNEL_PAIRS)
    with pytest.raises(L.LbmError, match="pairs kernel"):
        L.Lattice(128, 64, D, A, W, flags=L.KERNEL_PAIRS, n_gpus=2, device_ids=[0, 0])


def test_blown_up_lattice_reports_nan():
    nx, ny = 64, 12
    cells, obst = O.random_lattice(nx, ny, seed=2, p_obst=0.0, walls=False)
    obst[:] = 0
    cells[4:7, 19:22, :] = 0.0
    with L.Lattice(nx, ny, D, A, W, cells=cells, obstacles=obst, flags=L.KERNEL_PAIRS) as lat:
        av = lat.run(4)
    assert np.isnan(av).all(), av
