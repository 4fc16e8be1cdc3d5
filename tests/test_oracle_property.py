"""Property test (hypothesis): the oracle equals the compiled reference (strict build) bit
for bit on ANY small lattice -- shape, mask density, parameters -- for the fused step, and
the fused step equals the un-fused sequence of the five semantic functions."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle_lib as O


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
@given(nx=st.integers(2, 70), ny=st.integers(2, 24), seed=st.integers(0, 10 ** 6),
       p_obst=st.sampled_from([0.0, 0.05, 0.5, 1.0]), density=st.sampled_from([0.1, 0.37, 2.0]),
       accel=st.sampled_from([0.005, 0.05, 1.1, 3.0]), omega=st.sampled_from([0.7, 1.0, 1.85]),
       walls=st.booleans())
def test_oracle_equals_reference_everywhere(nx, ny, seed, p_obst, density, accel, omega, walls):
    lib = O.reference_lib("f32_strict")
    if lib is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cells, obst = O.random_lattice(nx, ny, seed=seed, density=density, p_obst=p_obst, walls=walls)
    if p_obst == 1.0:
        obst[:] = 1
    ra, rb, rav = O.ref_timestep_new2(lib, cells, obst, density, accel, omega)
    oa, ob, oav = O.timestep(cells, obst, density, accel, omega)
    assert np.array_equal(ra.view(np.uint32), oa.view(np.uint32))
    assert np.array_equal(rb.view(np.uint32), ob.view(np.uint32))
    assert np.float32(rav) == np.float32(oav) or (np.isnan(rav) and np.isnan(oav))
    unfused, uav = O.timestep_unfused(cells, obst, density, accel, omega)
    assert np.array_equal(unfused.view(np.uint32), ob.view(np.uint32))
    assert np.float32(uav) == np.float32(oav) or (np.isnan(uav) and np.isnan(oav))
