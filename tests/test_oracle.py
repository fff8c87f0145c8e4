"""CPU tests: the C oracle against the committed golden vectors (tests/golden/), the
SURVEY Appendix B known answers, and the reference's own unit tests restated.

Reference tests restated here:
  src/utils/custom/util_fns.rs:16-32   clip (3 tests)
  src/spaces/discrete.rs:27-41         Discrete::contains (2 tests)
  src/utils/seeding.rs:33-39, :11-20   rand_random seed echo (test + doctest)
"""
import math

import numpy as np
import pytest

import oracle


def ulps(a, b):
    a, b = np.float64(a), np.float64(b)
    if a == b:
        return 0
    return abs(int(a.view(np.int64)) - int(b.view(np.int64))) if (a > 0) == (b > 0) else 1 << 62


# ---- the reference's own unit tests --------------------------------------

def test_clip_value_beyond_upper_bound_returns_upper_bound():
    assert oracle.lib().orc_clip_i64(2, 0, 1) == 1


def test_clip_value_below_lower_bound_returns_lower_bound():
    assert oracle.lib().orc_clip_i64(-1, 0, 1) == 0


def test_clip_value_between_bounds_returns_value():
    assert oracle.lib().orc_clip_i64(1, -1, 2) == 1


def test_clip_f64_branch_order():
    L = oracle.lib()
    assert L.orc_clip(0.08, -0.07, 0.07) == 0.07
    assert L.orc_clip(-0.08, -0.07, 0.07) == -0.07
    assert L.orc_clip(0.01, -0.07, 0.07) == 0.01
    assert L.orc_clip(0.07, -0.07, 0.07) == 0.07


def test_discrete_contains_value_ge_upper_bound_false():
    L = oracle.lib()
    assert not L.orc_discrete_contains(3, 3)
    assert not L.orc_discrete_contains(3, 4)


def test_discrete_contains_value_lt_upper_bound_true():
    L = oracle.lib()
    assert L.orc_discrete_contains(3, 1)
    assert L.orc_discrete_contains(3, 2)


def test_rand_random_echoes_seed():
    L = oracle.lib()
    assert L.orc_rand_random(1, 42) == 42
    assert L.orc_rand_random(1, 64) == 64
    a, b = L.orc_rand_random(0, 0), L.orc_rand_random(0, 0)
    assert a != b  # OS entropy


# ---- Philox ---------------------------------------------------------------

def test_philox_known_answers(golden):
    for v in golden["philox4x32_10"]:
        assert oracle.philox4x32_10(v["ctr"], v["key"]) == v["out"]


# ---- defaults / spaces ----------------------------------------------------

def test_cartpole_defaults_and_spaces():
    p = oracle.default_params(oracle.CARTPOLE)
    assert (p.gravity, p.masscart, p.masspole, p.length, p.force_mag, p.tau) == \
        (9.8, 1.0, 0.1, 0.5, 10.0, 0.02)
    assert p.theta_threshold_radians == 0.20943951023931953
    assert p.x_threshold == 2.4 and p.kinematics_integrator == 0
    lo = (oracle.C.c_double * 4)()
    hi = (oracle.C.c_double * 4)()
    oracle.lib().orc_cartpole_observation_space(oracle.C.byref(p), lo, hi)
    assert list(hi) == [4.8, math.inf, 0.41887902047863906, math.inf]
    assert list(lo) == [-4.8, -math.inf, -0.41887902047863906, -math.inf]


def test_mountain_car_defaults_and_spaces():
    p = oracle.default_params(oracle.MOUNTAIN_CAR)
    assert (p.min_position, p.max_position, p.max_speed, p.goal_position, p.goal_velocity,
            p.force, p.gravity) == (-1.2, 0.6, 0.07, 0.5, 0.0, 0.001, 0.0025)
    lo = (oracle.C.c_double * 2)()
    hi = (oracle.C.c_double * 2)()
    oracle.lib().orc_mountain_car_observation_space(oracle.C.byref(p), lo, hi)
    assert list(lo) == [-1.2, -0.07] and list(hi) == [0.6, 0.07]


# ---- step known answers ---------------------------------------------------

APPENDIX_B_CARTPOLE = [
    ((0, 0, 0, 0), 1, (0.0, 0.3414634146341463, 0.0, -0.2926829268292683), False),
    ((0, 0, 0, 0), 0, (0.0, -0.3414634146341463, 0.0, 0.2926829268292683), False),
    ((0.01, -0.02, 0.03, 0.04), 1,
     (0.009600000000000001, 0.3161507699639987, 0.030799999999999998, -0.2430694901285738), False),
    ((0.01, -0.02, 0.03, 0.04), 0,
     (0.009600000000000001, -0.36646778424013676, 0.030799999999999998, 0.3419944516074735), False),
    ((2.39, 1.0, 0, 0), 1, (2.41, 1.3414634146341462, 0.0, -0.2926829268292683), True),
    ((0, 0, 0.2, 1.5), 0, (0.0, -0.35915586015046186, 0.23, 1.8408535750337613), True),
    ((-1.0, -2.0, -0.15, -1.0), 1,
     (-1.04, -1.639996108139277, -0.16999999999999998, -1.3334063584083258), False),
]

APPENDIX_B_MOUNTAIN_CAR = [
    ((-0.5, 0), 0, (-0.5011768430041692, -0.0011768430041692573), False),
    ((-0.5, 0), 1, (-0.5001768430041692, -0.00017684300416925727), False),
    ((-0.5, 0), 2, (-0.49917684300416926, 0.0008231569958307428), False),
    ((-1.2, -0.05), 0, (-1.2, 0.0), False),
    ((-1.19, -0.07), 0, (-1.2, 0.0), False),
    ((0.49, 0.07), 2, (0.56, 0.07), True),
    ((0.6, 0.07), 2, (0.6, 0.07), True),
    ((0.45, 0.04), 1, (0.4894524832822674, 0.0394524832822674), False),
]


@pytest.mark.parametrize("s,a,exp,done", APPENDIX_B_CARTPOLE)
def test_cartpole_appendix_b(s, a, exp, done):
    r = oracle.step_batch(oracle.CARTPOLE, np.array(s, dtype=np.float64)[:, None], [a])
    for got, want in zip(r["state"].ravel(), exp):
        assert ulps(got, want) <= 2, (got, want)
    assert bool(r["done"][0]) == done
    assert r["reward"][0] == 1.0


@pytest.mark.parametrize("s,a,exp,done", APPENDIX_B_MOUNTAIN_CAR)
def test_mountain_car_appendix_b(s, a, exp, done):
    r = oracle.step_batch(oracle.MOUNTAIN_CAR, np.array(s, dtype=np.float64)[:, None], [a])
    for got, want in zip(r["state"].ravel(), exp):
        assert ulps(got, want) <= 2, (got, want)
    assert bool(r["done"][0]) == done
    assert r["reward"][0] == -1.0


def _check_vectors(kind, vecs, params=None):
    st = np.array([v["state"] for v in vecs], dtype=np.float64).T
    act = [v["action"] for v in vecs]
    r = oracle.step_batch(kind, st, act, params=params)
    assert r["invalid"] == 0
    for i, v in enumerate(vecs):
        for k in range(st.shape[0]):
            got = r["state"][k, i]
            assert ulps(got, v["next_state"][k]) <= 2, (i, k, got, v["next_state"][k])
            mpv = v["next_state_mp"][k]
            assert abs(got - mpv) <= 1e-13 * max(1.0, abs(mpv)), (i, k, got, mpv)
        if "done" in v:
            assert bool(r["done"][i]) == v["done"], i
    return r


def test_cartpole_golden(golden):
    _check_vectors(oracle.CARTPOLE, golden["cartpole"])


def test_cartpole_semi_implicit_golden(golden):
    p = oracle.default_params(oracle.CARTPOLE)
    p.kinematics_integrator = 1
    _check_vectors(oracle.CARTPOLE, golden["cartpole_semi_implicit"], params=p)


def test_mountain_car_golden(golden):
    r = _check_vectors(oracle.MOUNTAIN_CAR, golden["mountain_car"])
    assert np.all(r["reward"] == -1.0)


def test_pendulum_golden(golden):
    vecs = golden["pendulum"]
    r = _check_vectors(oracle.PENDULUM, vecs)
    for i, v in enumerate(vecs):
        for k in range(3):
            assert abs(r["obs"][k, i] - v["obs_mp"][k]) <= 1e-13, (i, k)
            assert ulps(r["obs"][k, i], v["obs"][k]) <= 2
        assert abs(r["reward"][i] - v["reward_mp"]) <= 1e-12 * max(1.0, abs(v["reward_mp"]))
        assert not r["done"][i]


def test_cartpole_reward_sequence_after_termination(golden):
    """cartpole.rs:455-464: 1.0 while alive, 1.0 on the first done step, then 0.0;
    the state keeps integrating after done."""
    seq = golden["cartpole_reward_sequence"]
    st = np.zeros((4, 1))
    sbt = np.array([-1], dtype=np.int64)
    seen_done = False
    for i, v in enumerate(seq):
        r = oracle.step_batch(oracle.CARTPOLE, st, [1], sbt=sbt)
        st, sbt = r["state"], r["sbt"]
        assert bool(r["done"][0]) == v["done"], i
        assert r["reward"][0] == v["reward"], i
        for k in range(4):
            assert ulps(st[k, 0], v["state"][k]) <= 4 * (i + 1)
        seen_done |= v["done"]
    assert seen_done and seq[-1]["reward"] == 0.0


def test_invalid_actions_are_rejected():
    st = np.zeros((4, 3))
    r = oracle.step_batch(oracle.CARTPOLE, st, [0, 2, -1])
    assert r["invalid"] == 2
    st = np.zeros((2, 4))
    r = oracle.step_batch(oracle.MOUNTAIN_CAR, st, [0, 2, 3, -5])
    assert r["invalid"] == 2


def test_threshold_is_strict():
    # x lands exactly on 2.4 -> strict '>' -> not done (cartpole.rs:450-453)
    st = np.array([[2.4], [0.0], [0.0], [0.0]])
    r = oracle.step_batch(oracle.CARTPOLE, st, [1])
    assert r["state"][0, 0] == 2.4 and not r["done"][0]


# ---- reset ----------------------------------------------------------------

def test_reset_ranges_and_determinism():
    n = 20000
    a = oracle.reset_batch(oracle.CARTPOLE, n, seed=7)
    b = oracle.reset_batch(oracle.CARTPOLE, n, seed=7)
    c = oracle.reset_batch(oracle.CARTPOLE, n, seed=8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= -0.05 and a.max() < 0.05
    assert abs(a.mean()) < 2e-3 and abs(a.std() - 0.1 / math.sqrt(12)) < 1e-3
    m = oracle.reset_batch(oracle.MOUNTAIN_CAR, n, seed=7)
    assert m[0].min() >= -0.6 and m[0].max() < -0.4 and np.all(m[1] == 0.0)
    p = oracle.reset_batch(oracle.PENDULUM, n, seed=7)
    assert p[0].min() >= -math.pi and p[0].max() < math.pi
    assert p[1].min() >= -1.0 and p[1].max() < 1.0


def test_reset_is_sharding_invariant():
    full = oracle.reset_batch(oracle.CARTPOLE, 1000, seed=3)
    lo = oracle.reset_batch(oracle.CARTPOLE, 500, seed=3, global_env_offset=0)
    hi = oracle.reset_batch(oracle.CARTPOLE, 500, seed=3, global_env_offset=500)
    assert np.array_equal(full, np.concatenate([lo, hi], axis=1))


def test_reset_custom_bounds_and_mask():
    st0 = np.full((4, 8), 9.0)
    mask = np.array([1, 0, 1, 0, 1, 0, 1, 0], dtype=np.uint8)
    st = oracle.reset_batch(oracle.CARTPOLE, 8, seed=1, low=[1, 2, 3, 4], high=[2, 3, 4, 5],
                            mask=mask, state=st0)
    assert np.all(st[:, 1::2] == 9.0)
    for k in range(4):
        assert np.all((st[k, ::2] >= k + 1) & (st[k, ::2] < k + 2))


def test_bench_rollout_runs():
    t, cs = oracle.bench_rollout(oracle.CARTPOLE, 4096, 20, 2, 2, seed=0)
    assert t > 0 and math.isfinite(cs)
    t, cs = oracle.bench_rollout(oracle.MOUNTAIN_CAR, 4096, 20, 2, 2, seed=0)
    assert t > 0 and cs < 0
    t, cs = oracle.bench_rollout(oracle.PENDULUM, 4096, 20, 2, 1, seed=0)
    assert t > 0 and math.isfinite(cs)


def test_bench_regions_times_every_region_over_one_set_of_envs():
    # the reference arm of bench.py: burn-in, then regions of { warmup, timed steps }
    for kind, threads in ((oracle.CARTPOLE, 2), (oracle.MOUNTAIN_CAR, 3), (oracle.PENDULUM, 1)):
        times = oracle.bench_regions(kind, 4096, 5, 2, 7, 4, threads, seed=0)
        assert len(times) == 4 and all(0.0 < t < 5.0 for t in times), times



# ---- OrderedFloat total-order semantics (O64 = OrderedFloat<f64>, types.rs:4) ------------

def test_nan_follows_ordered_float_total_order():
    """The reference compares OrderedFloat values: NaN == NaN and NaN is greater than everything.
    So clip(NaN) is the RIGHT bound (util_fns.rs:2-10) and a NaN cart position / pole angle is
    'greater than the threshold', i.e. done (cartpole.rs:450-453)."""
    L = oracle.lib()
    nan = float("nan")
    assert L.orc_clip(nan, -0.07, 0.07) == 0.07
    assert L.orc_clip(math.inf, -0.07, 0.07) == 0.07 and L.orc_clip(-math.inf, -0.07, 0.07) == -0.07
    st = np.array([[nan, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, nan, 0.0], [0.0, 0.0, 0.0]])
    r = oracle.step_batch(oracle.CARTPOLE, st, [1, 1, 1])
    assert list(r["done"]) == [1, 1, 0]
    # MountainCar: a NaN position makes cos() NaN -> velocity clipped to +max_speed, position to max_position
    r = oracle.step_batch(oracle.MOUNTAIN_CAR, np.array([[nan], [0.0]]), [1])
    assert r["state"][0, 0] == 0.6 and r["state"][1, 0] == 0.07 and r["done"][0] == 1
