"""Property tests of the oracle (CPU, hypothesis): invariants the reference's formulas imply,
independent of any golden number.  The GPU counterpart at full size is
tests/test_parity_gpu.py::test_mirror_symmetry_at_full_size."""
import math

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle

finite = st.floats(allow_nan=False, allow_infinity=False, width=64, min_value=-1e6, max_value=1e6)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(v=st.floats(allow_nan=True, allow_infinity=True), lo=finite, width=st.floats(min_value=0, max_value=1e6))
def test_clip_properties(v, lo, width):
    """util_fns.rs:2-10 on OrderedFloat: result in [lo, hi]; identity inside; idempotent; NaN -> hi."""
    hi = lo + width
    L = oracle.lib()
    r = L.orc_clip(v, lo, hi)
    assert lo <= r <= hi
    if lo <= v <= hi:
        assert r == v
    if math.isnan(v):
        assert r == hi
    assert L.orc_clip(r, lo, hi) == r


@settings(max_examples=200, deadline=None, derandomize=True)
@given(n=st.integers(min_value=0, max_value=2 ** 40), v=st.integers(min_value=0, max_value=2 ** 41))
def test_discrete_contains_is_less_than(n, v):
    assert bool(oracle.lib().orc_discrete_contains(n, v)) == (v < n)   # discrete.rs:14-20


@settings(max_examples=200, deadline=None, derandomize=True)
@given(x=st.floats(-3, 3), xd=st.floats(-5, 5), th=st.floats(-0.5, 0.5), thd=st.floats(-5, 5), a=st.integers(0, 1))
def test_cartpole_mirror_symmetry(x, xd, th, thd, a):
    """The dynamics are odd under (state, force) -> (-state, -force): every IEEE operation involved
    is sign-symmetric and sin is odd / cos even in libm, so the symmetry is EXACT, not approximate."""
    s = np.array([[x], [xd], [th], [thd]])
    r1 = oracle.step_batch(oracle.CARTPOLE, s, [a])
    r2 = oracle.step_batch(oracle.CARTPOLE, -s, [1 - a])
    assert np.array_equal(r1["state"], -r2["state"])
    assert r1["done"][0] == r2["done"][0] and r1["reward"][0] == r2["reward"][0]


@settings(max_examples=300, deadline=None, derandomize=True)
@given(p=st.floats(-2, 1), v=st.floats(-0.2, 0.2), a=st.integers(0, 2))
def test_mountain_car_stays_in_its_box(p, v, a):
    """After the two clips (mountain_car.rs:413-416) the state is inside the observation space, the
    wall rule zeroes only negative velocities at the wall, and done implies the goal conditions."""
    r = oracle.step_batch(oracle.MOUNTAIN_CAR, np.array([[p], [v]]), [a])
    np_, nv = r["state"][0, 0], r["state"][1, 0]
    assert -1.2 <= np_ <= 0.6 and -0.07 <= nv <= 0.07
    if np_ == -1.2:
        assert nv >= 0.0
    assert bool(r["done"][0]) == (np_ >= 0.5 and nv >= 0.0)
    assert r["reward"][0] == -1.0


@settings(max_examples=200, deadline=None, derandomize=True)
@given(th=st.floats(-20, 20), thd=st.floats(-8, 8), u=st.floats(-5, 5))
def test_pendulum_invariants(th, thd, u):
    r = oracle.step_batch(oracle.PENDULUM, np.array([[th], [thd]]), [u])
    c, s, w = r["obs"][:, 0]
    assert abs(c * c + s * s - 1.0) < 1e-12 and abs(w) <= 8.0
    assert r["reward"][0] <= 0.0 and r["reward"][0] >= -(math.pi ** 2 + 0.1 * 64 + 0.001 * 4) - 1e-9
    # torque beyond the limit is clipped: same transition as the limit itself
    if abs(u) > 2.0:
        r2 = oracle.step_batch(oracle.PENDULUM, np.array([[th], [thd]]), [math.copysign(2.0, u)])
        assert np.array_equal(r["obs"], r2["obs"]) and r["reward"][0] == r2["reward"][0]
    # the observation is invariant under a 2 pi shift of theta (up to the rounding of the shift)
    r3 = oracle.step_batch(oracle.PENDULUM, np.array([[th + 2 * math.pi], [thd]]), [u])
    assert np.abs(r3["obs"] - r["obs"]).max() < 1e-9


@settings(max_examples=100, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 64 - 1), gid=st.integers(0, 2 ** 40), epoch=st.integers(0, 2 ** 32))
def test_reset_is_a_pure_function_of_seed_id_epoch(seed, gid, epoch):
    a = oracle.reset_batch(oracle.CARTPOLE, 1, seed=seed, global_env_offset=gid, epoch=epoch)
    b = oracle.reset_batch(oracle.CARTPOLE, 1, seed=seed, global_env_offset=gid, epoch=epoch)
    c = oracle.reset_batch(oracle.CARTPOLE, 3, seed=seed, global_env_offset=max(gid - 1, 0), epoch=epoch)
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, 0], c[:, 1 if gid >= 1 else 0])   # keyed by the GLOBAL id, not the local index
    assert np.all(a >= -0.05) and np.all(a < 0.05)
