"""Property test (hypothesis): for ANY small grid, mask density, physical parameters,
kernel variant, slab count and split of the run into calls, the strict CUDA path gives the
oracle's lattice bit for bit and the oracle's av_vels; and the device checksum equals the
numpy restatement on the downloaded lattice."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import lbm_b200 as L
import oracle_lib as O

pytestmark = pytest.mark.gpu

KERNELS = [L.KERNEL_VEC4, L.KERNEL_SCALAR, L.KERNEL_PERSISTENT, L.KERNEL_TMA, L.KERNEL_CLUSTER]


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large],
          derandomize=True)
@given(nx=st.integers(1, 300), ny=st.integers(2, 40), steps=st.integers(1, 7), seed=st.integers(0, 10 ** 6),
       p_obst=st.sampled_from([0.0, 0.02, 0.2, 0.6]), kernel=st.sampled_from(KERNELS),
       slabs=st.integers(1, 4), density=st.sampled_from([0.1, 0.37]), accel=st.sampled_from([0.005, 0.05, 1.1]),
       omega=st.sampled_from([0.7, 1.0, 1.85]), split=st.integers(0, 7), walls=st.booleans())
def test_any_configuration_matches_the_oracle(nx, ny, steps, seed, p_obst, kernel, slabs, density, accel, omega,
                                              split, walls):
    cells, obst = O.random_lattice(nx, ny, seed=seed, density=density, p_obst=p_obst, walls=walls)
    ref, _, av_ref = O.run(cells, obst, steps, density, accel, omega)
    n = 1 if kernel in (L.KERNEL_PERSISTENT, L.KERNEL_CLUSTER) else min(slabs, ny)
    first = min(split, steps)
    with L.Lattice(nx, ny, density, accel, omega, cells=cells, obstacles=obst, flags=L.STRICT | kernel, n_gpus=n,
                   device_ids=[0] * n) as lat:
        av = np.concatenate([lat.run(first), lat.run(steps - first)])
        got = lat.download()
        _, cs = lat.digest()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert cs == L.lattice_checksum(ref)
    ok = np.isfinite(av_ref)
    assert np.array_equal(np.isfinite(av), ok)
    np.testing.assert_allclose(av[ok].astype(np.float64), av_ref[ok], rtol=1e-6)
