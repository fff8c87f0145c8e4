"""GPU parity tests: the CUDA step path, called through the C ABI, against the f64 CPU oracle
on identical inputs.

Tolerance ("within 1e-6", BASELINE.json north_star):
    |gpu - ref| <= 1e-6 * max(1, |ref|)        for every float output.
done flags are compared wherever the f64 value is farther than 1e-6 from a termination
threshold; inside that band f32 and f64 may legitimately disagree and are only counted.
Integer / byte outputs (done outside the band, reward class, steps_beyond_terminated,
truncated) and everything that is GPU-vs-GPU (rollout vs single steps, vec widths, sharding,
host path vs device path) are compared bit-exactly.
"""
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

TOL = 1e-6
N_FULL = 1 << 20


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def g():
    import gym_rs_b200
    return gym_rs_b200


def mixed_err(gpu, ref):
    gpu = np.asarray(gpu, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(gpu - ref) / np.maximum(1.0, np.abs(ref))


def assert_within(gpu, ref, what, tol=TOL):
    e = mixed_err(gpu, ref)
    assert np.all(np.isfinite(np.asarray(gpu, dtype=np.float64))), what
    assert e.max() <= tol, f"{what}: max mixed err {e.max():.3e} at {np.unravel_index(e.argmax(), e.shape)}"
    return float(e.max())


def dev_actions(torch, a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def cartpole_inputs(n, seed=0):
    r = np.random.default_rng(seed)
    st = np.stack([r.uniform(-2.4, 2.4, n), r.uniform(-3, 3, n), r.uniform(-0.21, 0.21, n),
                   r.uniform(-3, 3, n)]).astype(np.float32)
    act = r.integers(0, 2, n).astype(np.int32)
    return st, act


def mountain_car_inputs(n, seed=0):
    r = np.random.default_rng(seed)
    st = np.stack([r.uniform(-1.2, 0.6, n), r.uniform(-0.07, 0.07, n)]).astype(np.float32)
    act = r.integers(0, 3, n).astype(np.int32)
    return st, act


def pendulum_inputs(n, seed=0):
    r = np.random.default_rng(seed)
    st = np.stack([r.uniform(-math.pi, math.pi, n), r.uniform(-8, 8, n)]).astype(np.float32)
    act = r.uniform(-2.5, 2.5, n).astype(np.float32)
    return st, act


def cartpole_band(ref_state, p=None):
    x, th = ref_state[0], ref_state[2]
    thr_x, thr_th = 2.4, 12 * 2 * math.pi / 360
    return (np.abs(np.abs(x) - thr_x) < 1e-6) | (np.abs(np.abs(th) - thr_th) < 1e-6)


# --------------------------------------------------------------------------------------
# single-step parity on 2^20 random (state, action) pairs
# --------------------------------------------------------------------------------------

def test_cartpole_single_step_parity_1m(torch, g):
    st, act = cartpole_inputs(N_FULL)
    env = g.CartPoleEnv(num_envs=N_FULL)
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    ref = oracle.step_batch(oracle.CARTPOLE, st, act, sbt=np.full(N_FULL, -1))
    got = out.observation.cpu().numpy()
    worst = assert_within(got, ref["state"], "cartpole state")
    assert np.array_equal(env.get_state(), got)  # observation IS the state
    band = cartpole_band(ref["state"])
    gd = out.done.cpu().numpy()
    assert np.array_equal(gd[~band], ref["done"][~band])
    assert band.sum() < 100
    assert np.array_equal(out.reward.cpu().numpy(), ref["reward"].astype(np.float32))
    assert not out.truncated.cpu().numpy().any()
    sbt = env.steps_beyond_terminated.cpu().numpy()
    assert np.array_equal(sbt[~band], ref["sbt"][~band])
    assert 0.02 < ref["done"].mean() < 0.9  # the inputs exercise both outcomes
    print(f"cartpole 1M single-step max mixed err {worst:.3e}, in-band {int(band.sum())}")
    env.close()


def test_mountain_car_single_step_parity_1m(torch, g):
    st, act = mountain_car_inputs(N_FULL)
    env = g.MountainCarEnv(num_envs=N_FULL)
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    ref = oracle.step_batch(oracle.MOUNTAIN_CAR, st, act)
    got = out.observation.cpu().numpy()
    worst = assert_within(got, ref["state"], "mountain car state")
    p, v = ref["state"]
    band = (np.abs(p - 0.5) < 1e-6) | (np.abs(v) < 1e-6) | (np.abs(p + 1.2) < 1e-6)
    assert np.array_equal(out.done.cpu().numpy()[~band], ref["done"][~band])
    assert np.all(out.reward.cpu().numpy() == -1.0)
    assert ref["done"].sum() > 1000
    print(f"mountain car 1M single-step max mixed err {worst:.3e}, in-band {int(band.sum())}")
    env.close()


def test_pendulum_single_step_parity_1m(torch, g):
    st, act = pendulum_inputs(N_FULL)
    env = g.PendulumEnv(num_envs=N_FULL)
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    ref = oracle.step_batch(oracle.PENDULUM, st, act)
    worst = assert_within(out.observation.cpu().numpy(), ref["obs"], "pendulum obs")
    assert_within(out.reward.cpu().numpy(), ref["reward"], "pendulum reward")
    gs = env.get_state()
    # stored theta is wrapped to [-pi, pi): compare modulo 2 pi
    d = gs[0].astype(np.float64) - ref["state"][0]
    d = d - 2 * math.pi * np.round(d / (2 * math.pi))
    th_err = np.abs(d) / np.maximum(1.0, np.abs(ref["state"][0]))   # the stated metric: 1e-6 * max(1, |ref|)
    assert th_err.max() <= TOL, th_err.max()
    assert np.abs(gs[0]).max() <= math.pi + 1e-6
    assert_within(gs[1], ref["state"][1], "pendulum theta_dot")
    assert not out.done.cpu().numpy().any()
    print(f"pendulum 1M single-step max mixed err {worst:.3e}")
    env.close()


def test_cartpole_semi_implicit_integrator(torch, g):
    n = 1 << 16
    st, act = cartpole_inputs(n, seed=3)
    env = g.CartPoleEnv(num_envs=n)
    p = env.params
    p.kinematics_integrator = 1
    env.params = p
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    op = oracle.default_params(oracle.CARTPOLE)
    op.kinematics_integrator = 1
    ref = oracle.step_batch(oracle.CARTPOLE, st, act, params=op)
    assert_within(out.observation.cpu().numpy(), ref["state"], "semi-implicit state")
    env.close()


def test_mutated_pub_fields_take_effect(torch, g):
    """The reference's constants are `pub` fields (cartpole.rs:63-80, mountain_car.rs:49-63)."""
    n = 1 << 14
    st, act = cartpole_inputs(n, seed=4)
    env = g.CartPoleEnv(num_envs=n)
    p = env.params
    p.gravity, p.masspole, p.length, p.force_mag, p.tau = 3.7, 0.25, 0.8, 7.0, 0.01
    p.x_threshold, p.theta_threshold_radians = 1.0, 0.1
    env.params = p
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    op = oracle.default_params(oracle.CARTPOLE)
    op.gravity, op.masspole, op.length, op.force_mag, op.tau = 3.7, 0.25, 0.8, 7.0, 0.01
    op.x_threshold, op.theta_threshold_radians = 1.0, 0.1
    ref = oracle.step_batch(oracle.CARTPOLE, st, act, params=op)
    assert_within(out.observation.cpu().numpy(), ref["state"], "mutated params state")
    x, th = ref["state"][0], ref["state"][2]
    band = (np.abs(np.abs(x) - 1.0) < 1e-6) | (np.abs(np.abs(th) - 0.1) < 1e-6)
    assert np.array_equal(out.done.cpu().numpy()[~band], ref["done"][~band])
    assert env.observation_space().high.x == 2.0
    env.close()

    st, act = mountain_car_inputs(n, seed=4)
    env = g.MountainCarEnv(num_envs=n)
    p = env.params
    p.force, p.gravity, p.max_speed, p.goal_position = 0.002, 0.003, 0.05, 0.4
    env.params = p
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    op = oracle.default_params(oracle.MOUNTAIN_CAR)
    op.force, op.gravity, op.max_speed, op.goal_position = 0.002, 0.003, 0.05, 0.4
    ref = oracle.step_batch(oracle.MOUNTAIN_CAR, st, act, params=op)
    assert_within(out.observation.cpu().numpy(), ref["state"], "mutated mountain car")
    env.close()


def test_special_values_follow_the_reference_semantics(torch, g):
    """Large angles (full-range sincosf path), huge velocities, infinities and NaN.  The reference
    compares OrderedFloat values (NaN greater than everything): a NaN cart/pole is done, and
    MountainCar clips a NaN to the RIGHT bound (util_fns.rs:2-10)."""
    nan, inf = float("nan"), float("inf")
    st = np.array([
        # x       x_dot   theta    theta_dot
        [0.0,     0.0,    1.0,     0.0],      # beyond pi/4: sincosf path
        [0.0,     0.0,    -3.0,    2.0],
        [1.0,     -1.0,   1.0e4,   0.5],      # needs the full range reduction
        [0.0,     0.0,    0.1,     1.0e3],    # theta_dot^2 = 1e6
        [1.0e6,   10.0,   0.0,     0.0],
        [nan,     0.0,    0.0,     0.0],
        [0.0,     0.0,    nan,     0.0],
        [inf,     0.0,    0.0,     0.0],
        [0.0,     nan,    0.0,     0.0],      # NaN velocity: x becomes NaN after the Euler update
    ], dtype=np.float32).T
    act = np.array([1, 0, 1, 0, 1, 1, 0, 1, 0], dtype=np.int32)
    env = g.CartPoleEnv(num_envs=st.shape[1])
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    ref = oracle.step_batch(oracle.CARTPOLE, st, act)
    got = out.observation.cpu().numpy()
    fin = np.isfinite(ref["state"])
    assert np.array_equal(np.isnan(got), np.isnan(ref["state"]))
    assert np.array_equal(np.isinf(got), np.isinf(ref["state"]))
    assert_within(np.where(fin, got, 0.0), np.where(fin, ref["state"], 0.0), "special values", tol=2e-6)
    assert np.array_equal(out.done.cpu().numpy(), ref["done"])
    assert ref["done"][5] == 1 and ref["done"][6] == 1 and ref["done"][8] == 1  # NaN is "greater than" the threshold
    env.close()

    st = np.array([[nan, 0.0], [0.0, nan], [inf, 0.0], [-inf, 0.0], [0.0, inf]], dtype=np.float32).T
    act = np.array([1, 1, 2, 0, 1], dtype=np.int32)
    env = g.MountainCarEnv(num_envs=st.shape[1])
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    ref = oracle.step_batch(oracle.MOUNTAIN_CAR, st, act)
    got = out.observation.cpu().numpy()
    assert np.isfinite(got).all() and np.isfinite(ref["state"]).all()
    assert_within(got, ref["state"], "mountain car special values")
    assert np.array_equal(out.done.cpu().numpy(), ref["done"])
    env.close()


# --------------------------------------------------------------------------------------
# golden vectors (tests/golden/step_vectors.json) through the GPU
# --------------------------------------------------------------------------------------

def test_golden_vectors_on_gpu(torch, g, golden):
    for key, cls, kind in (("cartpole", g.CartPoleEnv, oracle.CARTPOLE),
                           ("mountain_car", g.MountainCarEnv, oracle.MOUNTAIN_CAR)):
        vecs = golden[key]
        st = np.array([v["state"] for v in vecs], dtype=np.float64).T
        act = np.array([v["action"] for v in vecs], dtype=np.int32)
        env = cls(num_envs=len(vecs))
        env.set_state(st.astype(np.float32))
        out = env.step(dev_actions(torch, act))
        env.sync()
        got = out.observation.cpu().numpy()
        want = np.array([v["next_state_mp"] for v in vecs]).T
        # inputs were rounded to f32 first, so allow the propagated input rounding (<= 2e-7)
        assert_within(got, want, key, tol=1.5e-6)
        ref = oracle.step_batch(kind, st.astype(np.float32), act)
        assert_within(got, ref["state"], key + " (same f32 inputs)")
        env.close()
    vecs = golden["pendulum"]
    st = np.array([v["state"] for v in vecs], dtype=np.float32).T
    act = np.array([v["action"] for v in vecs], dtype=np.float32)
    env = g.PendulumEnv(num_envs=len(vecs))
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    env.sync()
    ref = oracle.step_batch(oracle.PENDULUM, st, act)
    assert_within(out.observation.cpu().numpy(), ref["obs"], "pendulum golden obs")
    assert_within(out.reward.cpu().numpy(), ref["reward"], "pendulum golden reward")
    env.close()


def test_appendix_b_exact_cases(torch, g):
    """Hand-checked boundary cases: termination by x / theta, wall rule, goal, clip."""
    env = g.CartPoleEnv(num_envs=4)
    st = np.array([[0, 0, 0, 0], [2.39, 1.0, 0, 0], [0, 0, 0.2, 1.5], [-2.39, -1.0, 0, 0]], dtype=np.float32).T
    env.set_state(st)
    out = env.step(dev_actions(torch, np.array([1, 1, 0, 0], dtype=np.int32)))
    env.sync()
    o = out.observation.cpu().numpy()
    assert abs(o[1, 0] - 0.3414634146341463) < 1e-6 and abs(o[3, 0] + 0.2926829268292683) < 1e-6
    assert list(out.done.cpu().numpy()) == [0, 1, 1, 1]
    assert list(out.reward.cpu().numpy()) == [1.0, 1.0, 1.0, 1.0]
    env.close()

    env = g.MountainCarEnv(num_envs=5)
    st = np.array([[-1.2, -0.05], [-1.19, -0.07], [0.49, 0.07], [0.6, 0.07], [-0.5, 0.0]], dtype=np.float32).T
    env.set_state(st)
    out = env.step(dev_actions(torch, np.array([0, 0, 2, 2, 1], dtype=np.int32)))
    env.sync()
    o = out.observation.cpu().numpy()
    assert o[0, 0] == np.float32(-1.2) and o[1, 0] == 0.0   # wall rule, mountain_car.rs:418-420
    assert o[0, 1] == np.float32(-1.2) and o[1, 1] == 0.0   # clip then wall rule
    assert abs(o[0, 2] - 0.56) < 1e-6 and o[1, 2] == np.float32(0.07)
    assert o[0, 3] == np.float32(0.6)                        # clipped to max_position
    assert abs(o[0, 4] + 0.5001768430041692) < 1e-6
    assert list(out.done.cpu().numpy()) == [0, 0, 1, 1, 0]
    env.close()


# --------------------------------------------------------------------------------------
# closed loop with per-step resync: the oracle is fed the device's f32 state each step
# --------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["cartpole", "mountain_car", "pendulum"])
def test_closed_loop_trajectory_parity(torch, g, kind):
    n, steps = 1 << 15, 500  # 500-step closed loop with per-step resync
    r = np.random.default_rng(11)
    if kind == "cartpole":
        env, ok = g.CartPoleEnv(num_envs=n), oracle.CARTPOLE
    elif kind == "mountain_car":
        env, ok = g.MountainCarEnv(num_envs=n), oracle.MOUNTAIN_CAR
    else:
        env, ok = g.PendulumEnv(num_envs=n), oracle.PENDULUM
    env.reset(seed=5)
    worst = 0.0
    sbt = np.full(n, -1, dtype=np.int64)
    disagreements = 0
    for t in range(steps):
        st = env.get_state()
        if kind == "pendulum":
            act = r.uniform(-2, 2, n).astype(np.float32)
        else:
            act = r.integers(0, 2 if kind == "cartpole" else 3, n).astype(np.int32)
        out = env.step(dev_actions(torch, act))
        env.sync()
        ref = oracle.step_batch(ok, st, act, sbt=sbt if kind == "cartpole" else None)
        if kind == "pendulum":
            worst = max(worst, assert_within(out.observation.cpu().numpy(), ref["obs"], f"{kind} t={t}"))
            assert_within(out.reward.cpu().numpy(), ref["reward"], f"{kind} reward t={t}")
        else:
            worst = max(worst, assert_within(out.observation.cpu().numpy(), ref["state"], f"{kind} t={t}"))
            gd = out.done.cpu().numpy()
            diff = gd != ref["done"]
            if kind == "cartpole":
                assert not (diff & ~cartpole_band(ref["state"])).any()
                gs = env.steps_beyond_terminated.cpu().numpy()
                sbt = gs.astype(np.int64)  # resync
                rr = out.reward.cpu().numpy()
                assert np.array_equal(rr[~diff], ref["reward"][~diff].astype(np.float32))
            disagreements += int(diff.sum())
    assert disagreements < 50
    if kind == "cartpole":
        # without reset practically every pole has fallen: steps_beyond_terminated is Some(_)
        assert (sbt >= 0).mean() > 0.99
    print(f"{kind} closed loop {steps} steps: max mixed err {worst:.3e}, in-band done flips {disagreements}")
    env.close()


def test_cartpole_reward_sequence_after_termination(torch, g, golden):
    """cartpole.rs:455-464 through the device: 1.0, ..., 1.0 (first done), 0.0, 0.0, ..."""
    seq = golden["cartpole_reward_sequence"]
    env = g.CartPoleEnv(num_envs=3)
    env.set_state(np.zeros((4, 3), dtype=np.float32))
    ones = dev_actions(torch, np.ones(3, dtype=np.int32))
    for i, v in enumerate(seq):
        out = env.step(ones)
        env.sync()
        assert bool(out.done.cpu().numpy()[0]) == v["done"], i
        assert float(out.reward.cpu().numpy()[0]) == v["reward"], i
        assert_within(out.observation.cpu().numpy()[:, 0], v["state"], f"step {i}", tol=1e-5)
    assert int(env.steps_beyond_terminated.cpu().numpy()[0]) > 0
    env.close()


# --------------------------------------------------------------------------------------
# reset
# --------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["cartpole", "mountain_car", "pendulum"])
def test_reset_matches_oracle_philox_stream(torch, g, kind):
    n = 100003
    cls, ok = {"cartpole": (g.CartPoleEnv, oracle.CARTPOLE), "mountain_car": (g.MountainCarEnv, oracle.MOUNTAIN_CAR),
               "pendulum": (g.PendulumEnv, oracle.PENDULUM)}[kind]
    env = cls(num_envs=n, global_env_offset=12345)
    obs, info = env.reset(seed=42, return_info=True)
    assert info == ()
    st = env.get_state()
    obs_at_reset = obs.cpu().numpy()  # obs is a live view of the device buffer: copy it now
    ref = oracle.reset_batch(ok, n, seed=42, global_env_offset=12345)
    assert np.abs(st - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())
    # a reset is a pure function of the seed, like re-seeding PCG64 in the reference
    env.reset(seed=42)
    assert np.array_equal(env.get_state(), st)
    env.reset(seed=43)
    assert not np.array_equal(env.get_state(), st)
    _, info = env.reset()
    assert info is None and env.rand_random()[1] not in (42, 43)
    if kind == "cartpole":
        assert st.min() >= -0.05 and st.max() < 0.05
        assert abs(st.mean()) < 1e-3 and abs(st.std() - 0.1 / math.sqrt(12)) < 1e-3
    elif kind == "mountain_car":
        assert st[0].min() >= -0.6 and st[0].max() < -0.4 and np.all(st[1] == 0)
    else:
        assert st[0].min() >= -math.pi - 1e-6 and st[0].max() <= math.pi and np.abs(st[1]).max() <= 1
        o = obs_at_reset
        assert_within(o[0], np.cos(st[0].astype(np.float64)), "reset cos")
        assert_within(o[1], np.sin(st[0].astype(np.float64)), "reset sin")
    env.close()


def test_reset_options_and_mask(torch, g):
    n = 4096
    env = g.CartPoleEnv(num_envs=n)
    lo = g.CartPoleObservation(1, 2, 3, 4)
    hi = g.CartPoleObservation(2, 3, 4, 5)
    env.reset(seed=9, options=g.BoxR(lo, hi))
    st = env.get_state()
    for k in range(4):
        assert st[k].min() >= k + 1 and st[k].max() < k + 2
    ref = oracle.reset_batch(oracle.CARTPOLE, n, seed=9, low=[1, 2, 3, 4], high=[2, 3, 4, 5])
    assert np.abs(st - ref).max() < 1e-6
    mask = np.zeros(n, dtype=np.uint8)
    mask[::3] = 1
    env.reset(seed=10, mask=torch.as_tensor(mask).cuda())
    st2 = env.get_state()
    assert np.array_equal(st2[:, mask == 0], st[:, mask == 0])
    assert np.abs(st2[:, mask == 1]).max() < 0.05  # options apply to one call only
    env.close()
    # mountain car ignores velocity bounds (mountain_car.rs:162-167)
    mc = g.MountainCarEnv(num_envs=n)
    mc.reset(seed=1, options=g.BoxR(g.MountainCarObservation(0.1, 0.01), g.MountainCarObservation(0.2, 0.05)))
    st = mc.get_state()
    assert st[0].min() >= 0.1 and st[0].max() < 0.2 and np.all(st[1] == 0)
    mc.close()


def test_autoreset_resamples_done_envs_like_the_oracle(torch, g):
    n = 1 << 16
    st, act = cartpole_inputs(n, seed=21)
    env = g.CartPoleEnv(num_envs=n, global_env_offset=1000)
    env.reset(seed=77)
    env.set_state(st)
    out = env.step(dev_actions(torch, act), autoreset=True)
    env.sync()
    ref = oracle.step_batch(oracle.CARTPOLE, st, act)
    gd = out.done.cpu().numpy().astype(bool)
    band = cartpole_band(ref["state"])
    assert np.array_equal(gd[~band], ref["done"][~band].astype(bool))
    got = out.observation.cpu().numpy()
    assert_within(got[:, ~gd], ref["state"][:, ~gd], "surviving envs")
    # done envs hold a fresh state: Philox(seed, global id, epoch = step index + 1)
    fresh = oracle.reset_batch(oracle.CARTPOLE, n, seed=77, global_env_offset=1000, epoch=1)
    assert np.abs(got[:, gd] - fresh[:, gd]).max() <= 1e-7
    assert np.all(out.reward.cpu().numpy() == 1.0)
    assert (env.steps_beyond_terminated.cpu().numpy() == -1).all()
    # the next step draws from the next epoch
    st2 = env.get_state()
    out = env.step(dev_actions(torch, act), autoreset=True)
    env.sync()
    gd2 = out.done.cpu().numpy().astype(bool)
    fresh2 = oracle.reset_batch(oracle.CARTPOLE, n, seed=77, global_env_offset=1000, epoch=2)
    assert gd2.any()
    assert np.abs(out.observation.cpu().numpy()[:, gd2] - fresh2[:, gd2]).max() <= 1e-7
    ref2 = oracle.step_batch(oracle.CARTPOLE, st2, act)
    assert_within(out.observation.cpu().numpy()[:, ~gd2], ref2["state"][:, ~gd2], "second step")
    env.close()


def test_long_autoreset_rollout_stays_in_the_live_region(torch, g):
    """1M envs x 300 random-action steps with auto-reset: every state stays finite and inside
    the live region plus one step; episode statistics match the oracle's scalar loop."""
    n = N_FULL
    env = g.CartPoleEnv(num_envs=n)
    env.reset(seed=0)
    gen = torch.Generator(device="cuda").manual_seed(1)
    dones = 0
    steps = 300
    for t in range(steps):
        act = torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int32)
        out = env.step(act, autoreset=True)
        dones += int(out.done.sum().item())
    env.sync()
    st = env.get_state()
    assert np.isfinite(st).all()
    assert np.abs(st[0]).max() <= 2.4 + 1e-6 and np.abs(st[2]).max() <= 0.2095
    mean_len = n * steps / dones
    # reference dynamics under uniform random actions: mean episode length ~22 steps
    assert 18.0 < mean_len < 27.0, mean_len
    env.close()


# --------------------------------------------------------------------------------------
# GPU-vs-GPU invariants (bit-exact)
# --------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["cartpole", "mountain_car", "pendulum"])
def test_vec_widths_and_ragged_sizes_are_bit_identical(torch, g, kind):
    n = 100003  # not a multiple of 4: exercises the scalar tail
    cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[kind]
    r = np.random.default_rng(5)
    if kind == "pendulum":
        acts = [dev_actions(torch, r.uniform(-2, 2, n).astype(np.float32)) for _ in range(6)]
    else:
        acts = [dev_actions(torch, r.integers(0, 2 if kind == "cartpole" else 3, n).astype(np.int32))
                for _ in range(6)]
    results = []
    for vec, block in ((4, 256), (2, 128), (1, 64), (0, 0)):
        env = cls(num_envs=n, time_limit=True)
        env.set_launch_config(vec=vec, block=block, pdl=1)
        env.reset(seed=3)
        for a in acts:
            out = env.step(a, autoreset=True)
        env.sync()
        results.append((env.get_state(), out.observation.cpu().numpy(), out.reward.cpu().numpy(),
                        out.done.cpu().numpy(), out.truncated.cpu().numpy()))
        env.close()
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert np.array_equal(a, b)


def test_step_pass_brackets_a_multi_stream_pass_with_events(torch, g):
    """gymrs_step_pass = gymrs_step_many + a begin event (recorded on the first handle's stream, the other streams
    of the pass fork from it) + an end event (recorded there after the other streams have been joined)."""
    import ctypes as C
    from gym_rs_b200 import _capi
    L = _capi.load()
    n = 1 << 18
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    envs = [g.CartPoleEnv(num_envs=n, global_env_offset=i * n) for i in range(4)]
    refs = [g.CartPoleEnv(num_envs=n, global_env_offset=i * n) for i in range(4)]
    gen = torch.Generator(device="cuda").manual_seed(2)
    acts = [torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(4)]
    for i, (e, r) in enumerate(zip(envs, refs)):
        e.reset(seed=9)
        r.reset(seed=9)
        e.sync()
        e.set_stream((s1 if i % 2 == 0 else s2).cuda_stream)
    torch.cuda.synchronize()
    b, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record(s1)
    e_.record(s1)  # creates the CUDA events
    torch.cuda.synchronize()
    order = [0, 1, 2, 3, 0, 1, 2, 3]
    hs = (C.c_void_p * len(order))(*[envs[i].handle.value if hasattr(envs[i].handle, "value") else envs[i].handle for i in order])
    ap = (C.c_void_p * len(order))(*[acts[(k + i) % 4].data_ptr() for k, i in enumerate(order)])
    done = C.c_uint32(0)
    _capi.check(L.gymrs_step_pass(hs, ap, len(order), _capi.STEP_AUTORESET, C.c_void_p(b.cuda_event), C.c_void_p(e_.cuda_event),
                                  C.byref(done)))
    assert done.value == len(order)
    e_.synchronize()  # ONE event covers both streams
    assert b.elapsed_time(e_) > 0
    for k, i in enumerate(order):
        refs[i].step(acts[(k + i) % 4], autoreset=True)
    for e, r in zip(envs, refs):
        r.sync()
        # no sync on e: the end event alone must have made both streams' results visible
        assert torch.equal(e._t_state, r._t_state) and torch.equal(e._t_done, r._t_done)
    # without events it is gymrs_step_many; an empty pass with events is refused
    _capi.check(L.gymrs_step_pass(hs, ap, 2, _capi.STEP_AUTORESET, None, None, None))
    assert L.gymrs_step_pass(hs, ap, 0, _capi.STEP_AUTORESET, C.c_void_p(b.cuda_event), None, None) == _capi.ERR_BAD_ARG
    for x in envs + refs:
        x.close()


def test_billion_env_batch_indexing(torch, g):
    """Maximum-size end of the range: one handle with 2^30 + 1027 MountainCar instances (~20 GB, ragged tail).
    Env indices are 32-bit inside a launch, byte offsets are not: rows are 4 GB apart and slices around the
    2^29-th / 2^30-th env and the tail must step exactly like the oracle says, reset keyed by global env id."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~20 GB of free device memory")
    n = (1 << 30) + 1027
    env = g.MountainCarEnv(num_envs=n, global_env_offset=7)
    env.reset(seed=5)
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = torch.randint(0, 3, (n,), generator=gen, device="cuda", dtype=torch.int32)
    slices = [slice(0, 4096), slice((1 << 29) - 2048, (1 << 29) + 2048), slice((1 << 30) - 2048, n)]
    before = [env.state[:, sl].clone() for sl in slices]
    for sl, b in zip(slices, before):  # the reset stream at these global ids
        ref = oracle.reset_batch(oracle.MOUNTAIN_CAR, sl.stop - sl.start, seed=5, global_env_offset=7 + sl.start)
        assert np.abs(b.cpu().numpy() - ref).max() <= 1e-7
    out = env.step(acts, autoreset=False)
    env.sync()
    for sl, b in zip(slices, before):
        ref = oracle.step_batch(oracle.MOUNTAIN_CAR, b.cpu().numpy(), acts[sl].cpu().numpy())
        assert_within(out.observation[:, sl].cpu().numpy(), ref["state"], f"envs {sl.start}..{sl.stop}")
        assert np.all(out.reward[sl].cpu().numpy() == -1.0)
        assert np.array_equal(out.done[sl].cpu().numpy(), ref["done"])
    # every env moved by at most max_speed and stayed in the box (a whole-batch property, checked on the device)
    assert float(out.observation[0].min()) >= -1.2 - 1e-6 and float(out.observation[0].max()) <= 0.6 + 1e-6
    assert float(out.observation[1].abs().max()) <= 0.07 + 1e-7
    env.close()
    del acts, before, out
    torch.cuda.empty_cache()


@pytest.mark.parametrize("kind", ["cartpole", "mountain_car", "pendulum"])
@pytest.mark.parametrize("time_limit", [False, True])
def test_wide_occupancy_build_is_bit_identical(torch, g, kind, time_limit):
    """gymrs_set_launch_occupancy(1): the same step kernel under a tighter register budget (more resident
    CTAs).  Same bits as the default build: plain and chained launches, auto-reset, with and without
    the time limit, a ragged batch size, and a chain that switches between the two builds."""
    n = (1 << 18) + 1001
    cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[kind]
    r = np.random.default_rng(11)
    if kind == "pendulum":
        acts = [dev_actions(torch, r.uniform(-2, 2, n).astype(np.float32)) for _ in range(12)]
    else:
        acts = [dev_actions(torch, r.integers(0, 2 if kind == "cartpole" else 3, n).astype(np.int32))
                for _ in range(12)]
    results = []
    for wide, pdl in ((False, 1), (True, 1), (True, 2), ("alternate", 2)):
        env = cls(num_envs=n, time_limit=time_limit)
        env.set_launch_config(vec=0, block=0, pdl=pdl)
        env.reset(seed=3)
        for k, a in enumerate(acts):
            env.set_launch_occupancy(bool(k & 1) if wide == "alternate" else wide)
            out = env.step(a, autoreset=(k % 5 != 4))  # every fifth step leaves finished envs alone
        env.sync()
        results.append((env.get_state(), out.observation.cpu().numpy(), out.reward.cpu().numpy(),
                        out.done.cpu().numpy(), out.truncated.cpu().numpy()))
        env.close()
    assert results[0][3].any() or kind != "cartpole"
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert np.array_equal(a, b)
    with pytest.raises(Exception):
        e = cls(num_envs=8)
        try:
            e._L.gymrs_set_launch_occupancy.restype  # noqa: B018 (the symbol exists)
            from gym_rs_b200 import _capi
            _capi.check(e._L.gymrs_set_launch_occupancy(e._h, 2))
        finally:
            e.close()


@pytest.mark.parametrize("kind,n,vec,block", [("cartpole", N_FULL, 0, 0), ("cartpole", 300001, 1, 32),
                                              ("mountain_car", 1 << 19, 2, 64), ("pendulum", 777777, 4, 128),
                                              # vec = 8: the persistent TMA-staged kernel (ragged last tile included)
                                              ("cartpole", N_FULL + 516, 8, 0), ("pendulum", 1 << 18, 8, 0)])
def test_chained_pdl_launches_match_plain_launches(torch, g, kind, n, vec, block):
    """pdl = 2 pipelines back-to-back steps: a CTA waits only for the same-index CTA of the handle's
    previous step (per-CTA release/acquire flags) instead of the whole previous grid.  Results must
    be bit-identical to plain stream-ordered launches, also when several handles interleave on one
    stream and when other calls (reset, set_state, rollout, host step) break the chain."""
    cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[kind]
    gen = torch.Generator(device="cuda").manual_seed(7)
    if kind == "pendulum":
        acts = [torch.rand((n,), generator=gen, device="cuda") * 4 - 2 for _ in range(8)]
    else:
        acts = [torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(8)]
    torch.cuda.synchronize()
    finals = {}
    # (launch width, pdl): plain vs chained launches of the same kernel; for the TMA-staged kernel
    # (vec = 8) additionally the plain step_kernel, which must give the same bits
    configs = [(vec, 0), (vec, 2)] + ([(4, 1)] if vec == 8 else [])
    for cfg_vec, pdl in configs:
        envs = [cls(num_envs=n, global_env_offset=k * n) for k in range(3)]
        for e in envs:
            e.set_launch_config(vec=cfg_vec, block=block, pdl=pdl)
            e.reset(seed=11)
        for t in range(240):
            for k, e in enumerate(envs):  # three handles interleaved on one stream
                e.step(acts[(t + k) % 8], autoreset=True)
            if t == 100:   # chain breakers in the middle of the run
                st = envs[0].get_state()
                envs[0].set_state(st)
                envs[1].reset(seed=5)
                envs[2].rollout(torch.stack(acts[:2]), autoreset=True)
        for e in envs:
            e.sync()
        finals[(cfg_vec, pdl)] = [(e.get_state(), e._t_reward.cpu().numpy(), e._t_done.cpu().numpy()) for e in envs]
        for e in envs:
            e.close()
    base = finals[configs[0]]
    for cfg in configs[1:]:
        for a, b in zip(base, finals[cfg]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y), cfg


def test_unaligned_action_pointer_falls_back_to_scalar_lanes(torch, g):
    n = 8192
    st, act = cartpole_inputs(n + 1, seed=8)
    big = dev_actions(torch, act)
    env = g.CartPoleEnv(num_envs=n)
    env.set_state(st[:, 1:])
    out = env.step(big[1:], autoreset=False)  # 4-byte aligned only
    env.sync()
    ref = oracle.step_batch(oracle.CARTPOLE, st[:, 1:], act[1:])
    assert_within(out.observation.cpu().numpy(), ref["state"], "unaligned actions")
    env.close()


@pytest.mark.parametrize("kind", ["cartpole", "mountain_car", "pendulum"])
def test_rollout_equals_single_steps(torch, g, kind):
    n, k = 50001, 17
    cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[kind]
    r = np.random.default_rng(6)
    if kind == "pendulum":
        acts = dev_actions(torch, r.uniform(-2, 2, (k, n)).astype(np.float32))
    else:
        acts = dev_actions(torch, r.integers(0, 2, (k, n)).astype(np.int32))
    a = cls(num_envs=n)
    b = cls(num_envs=n)
    a.reset(seed=2)
    b.reset(seed=2)
    od = a.obs_dim
    obs_out = torch.zeros((k, od, n), device="cuda")
    rew_out = torch.zeros((k, n), device="cuda")
    done_out = torch.zeros((k, n), device="cuda", dtype=torch.uint8)
    a.rollout(acts, obs_out, rew_out, done_out, autoreset=True)
    a.sync()
    for t in range(k):
        out = b.step(acts[t], autoreset=True)
        assert torch.equal(out.observation, obs_out[t]), t
        assert torch.equal(out.reward, rew_out[t]), t
        assert torch.equal(out.done, done_out[t]), t
    b.sync()
    assert np.array_equal(a.get_state(), b.get_state())
    # and the two handles keep agreeing afterwards (same epoch counter)
    o1 = a.step(acts[0], autoreset=True).observation.clone()
    o2 = b.step(acts[0], autoreset=True).observation.clone()
    assert torch.equal(o1, o2)
    a.close()
    b.close()


def test_step_host_equals_device_step(torch, g):
    n = (1 << 19) + 777  # several pipeline chunks + ragged tail
    st, act = cartpole_inputs(n, seed=9)
    a = g.CartPoleEnv(num_envs=n)
    b = g.CartPoleEnv(num_envs=n)
    for e in (a, b):
        e.reset(seed=4)
        e.set_state(st)
    obs = torch.empty((4, n), dtype=torch.float32).pin_memory()
    rew = torch.empty(n, dtype=torch.float32).pin_memory()
    dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    tr = torch.empty(n, dtype=torch.uint8).pin_memory()
    hact = torch.as_tensor(act).pin_memory()
    for _ in range(3):
        a.step_host(hact, obs, rew, dn, tr, autoreset=True)
        out = b.step(dev_actions(torch, act), autoreset=True)
        b.sync()
        assert torch.equal(obs, out.observation.cpu())
        assert torch.equal(rew, out.reward.cpu())
        assert torch.equal(dn, out.done.cpu())
        assert not tr.any()
    # pageable numpy buffers work too
    o2 = np.empty((4, n), dtype=np.float32)
    r2 = np.empty(n, dtype=np.float32)
    d2 = np.empty(n, dtype=np.uint8)
    a.step_host(act, o2, r2, d2, None, autoreset=True)
    out = b.step(dev_actions(torch, act), autoreset=True)
    b.sync()
    assert np.array_equal(o2, out.observation.cpu().numpy())
    a.close()
    b.close()


def test_async_host_pipeline_equals_synchronous_host_steps(torch, g):
    """gymrs_step_host_async / gymrs_host_wait (double-buffered pinned delivery overlapped with the
    next step) must deliver exactly what the synchronous host step delivers, step by step, and
    other entry points must wait for in-flight host steps."""
    n = (1 << 20) + 4096
    a = g.MountainCarEnv(num_envs=n)
    b = g.MountainCarEnv(num_envs=n)
    for e in (a, b):
        e.reset(seed=8)
    r = np.random.default_rng(2)
    acts = [torch.as_tensor(r.integers(0, 3, n).astype(np.int32)).pin_memory() for _ in range(6)]
    bufs = [dict(obs=torch.empty((2, n), dtype=torch.float32).pin_memory(),
                 rew=torch.empty(n, dtype=torch.float32).pin_memory(),
                 done=torch.empty(n, dtype=torch.uint8).pin_memory()) for _ in range(2)]
    ref = dict(obs=torch.empty((2, n), dtype=torch.float32).pin_memory(),
               rew=torch.empty(n, dtype=torch.float32).pin_memory(),
               done=torch.empty(n, dtype=torch.uint8).pin_memory())
    tickets = []
    expected = []
    for t, act in enumerate(acts):
        s = bufs[t % 2]
        if t >= 2:  # this buffer set is about to be reused: its step must be finished and checked
            a.host_wait(tickets[t - 2])
            assert torch.equal(s["obs"], expected[t - 2][0]) and torch.equal(s["done"], expected[t - 2][2])
        tickets.append(a.step_host_async(act, s["obs"], s["rew"], s["done"], None, autoreset=True))
        b.step_host(act, ref["obs"], ref["rew"], ref["done"], None, autoreset=True)
        expected.append((ref["obs"].clone(), ref["rew"].clone(), ref["done"].clone()))
    for t in (len(acts) - 2, len(acts) - 1):
        a.host_wait(tickets[t])
        s = bufs[t % 2]
        assert torch.equal(s["obs"], expected[t][0])
        assert torch.equal(s["rew"], expected[t][1])
        assert torch.equal(s["done"], expected[t][2])
    # a device step right after an un-waited async host step is ordered after it
    tk = a.step_host_async(acts[0], bufs[0]["obs"], bufs[0]["rew"], bufs[0]["done"], None, autoreset=True)
    da = acts[1].cuda()
    out = a.step(da, autoreset=True)
    a.sync()
    a.host_wait(tk)
    b.step_host(acts[0], ref["obs"], ref["rew"], ref["done"], None, autoreset=True)
    assert torch.equal(bufs[0]["obs"], ref["obs"])
    ob = b.step(da, autoreset=True)
    b.sync()
    assert torch.equal(out.observation, ob.observation)
    a.close()
    b.close()


def test_mountain_car_autoreset_resamples_done_envs_like_the_oracle(torch, g):
    """BASELINE config 3 (MountainCar, auto-reset on done): envs that reach the goal are re-sampled in
    the same launch from Philox(seed, global id, epoch = step index + 1); everything else follows
    mountain_car.rs:398-435."""
    n = 1 << 16
    r = np.random.default_rng(31)
    # near the goal, moving right: a good share of the envs finish in one step
    st = np.stack([r.uniform(0.40, 0.60, n), r.uniform(-0.02, 0.07, n)]).astype(np.float32)
    act = r.integers(0, 3, n).astype(np.int32)
    env = g.MountainCarEnv(num_envs=n, global_env_offset=5000)
    env.reset(seed=123)
    env.set_state(st)
    for epoch in (1, 2):
        out = env.step(dev_actions(torch, act), autoreset=True)
        env.sync()
        ref = oracle.step_batch(oracle.MOUNTAIN_CAR, st, act)
        gd = out.done.cpu().numpy().astype(bool)
        rp, rv = ref["state"]
        band = (np.abs(rp - 0.5) < 1e-6) | (np.abs(rv) < 1e-9)
        assert np.array_equal(gd[~band], ref["done"][~band].astype(bool))
        assert gd.mean() > 0.05 if epoch == 1 else True
        got = out.observation.cpu().numpy()
        assert_within(got[:, ~gd], ref["state"][:, ~gd], f"surviving envs, step {epoch}")
        fresh = oracle.reset_batch(oracle.MOUNTAIN_CAR, n, seed=123, global_env_offset=5000, epoch=epoch)
        assert np.abs(got[0, gd] - fresh[0, gd]).max() <= 1e-7 if gd.any() else True
        assert np.all(got[1, gd] == 0.0)                                  # mountain_car.rs:162-167
        assert np.all(out.reward.cpu().numpy() == -1.0)                   # :423, terminal step included
        assert not out.truncated.cpu().numpy().any()                      # :432
        st = env.get_state()
        assert np.array_equal(st, got)
        # second round: push the fresh envs (which sit in the valley) and the rest once more
    env.close()


def test_pendulum_autoreset_on_truncation_matches_the_oracle(torch, g):
    """Pendulum never terminates, so its auto-reset fires through the TimeLimit only: at the horizon
    every env is re-sampled from Philox(seed, global id, epoch = step index + 1) and the returned
    observation is (cos, sin, theta_dot) of the FRESH state."""
    n = 1 << 16
    r = np.random.default_rng(32)
    env = g.PendulumEnv(num_envs=n, time_limit=True, global_env_offset=77)
    p = env.params
    p.max_episode_steps = 3
    env.params = p
    env.reset(seed=9)
    for t in range(1, 8):
        st = env.get_state()
        act = r.uniform(-2.5, 2.5, n).astype(np.float32)
        out = env.step(dev_actions(torch, act), autoreset=True)
        env.sync()
        ref = oracle.step_batch(oracle.PENDULUM, st, act)
        assert_within(out.reward.cpu().numpy(), ref["reward"], f"reward t={t}")  # the cost of the step that ended the episode
        tr = out.truncated.cpu().numpy().astype(bool)
        assert tr.all() == (t % 3 == 0) and tr.any() == (t % 3 == 0)
        assert not out.done.cpu().numpy().any()
        obs = out.observation.cpu().numpy()
        if t % 3 == 0:
            fresh = oracle.reset_batch(oracle.PENDULUM, n, seed=9, global_env_offset=77, epoch=t)
            gs = env.get_state()
            assert np.abs(gs - fresh).max() <= 5e-7
            want = np.stack([np.cos(fresh[0]), np.sin(fresh[0]), fresh[1]])
            assert_within(obs, want, f"fresh observation t={t}")
            assert (core_elapsed(env) == 0).all()
        else:
            assert_within(obs, ref["obs"], f"obs t={t}")
    env.close()


@pytest.mark.parametrize("kind", ["cartpole", "mountain_car", "pendulum"])
def test_rollout_host_equals_synchronous_host_steps(torch, g, kind):
    """gymrs_rollout_host (the pipelined host loop inside the library) delivers, step by step and in
    order, exactly what gymrs_step_host delivers -- also with the compact wire formats."""
    n = (1 << 19) + 3 * 1024 + 5
    cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[kind]
    r = np.random.default_rng(3)
    steps, S = 7, 3
    a, b, c = cls(num_envs=n, time_limit=True), cls(num_envs=n, time_limit=True), cls(num_envs=n, time_limit=True)
    for e in (a, b, c):
        pp = e.params
        pp.max_episode_steps = 4
        e.params = pp
        e.reset(seed=8)
    od = a.obs_dim
    if kind == "pendulum":
        acts = torch.as_tensor(r.uniform(-2, 2, (steps, n)).astype(np.float32)).pin_memory()
    else:
        acts = torch.as_tensor(r.integers(0, 2 if kind == "cartpole" else 3, (steps, n)).astype(np.int32)).pin_memory()
    obs = torch.empty((S, od, n), dtype=torch.float32).pin_memory()
    rew = torch.empty((S, n), dtype=torch.float32).pin_memory()
    done = torch.empty((S, n), dtype=torch.uint8).pin_memory()
    trunc = torch.empty((S, n), dtype=torch.uint8).pin_memory()
    seen = []

    def on_step(t, slot):
        assert slot == t % S
        seen.append((t, obs[slot].clone(), rew[slot].clone(), done[slot].clone(), trunc[slot].clone()))

    a.rollout_host(acts, obs, rew, done, trunc, autoreset=True, on_step=on_step)
    assert [t for t, *_ in seen] == list(range(steps))
    ro, rr, rd, rt = (torch.empty((od, n)), torch.empty(n), torch.empty(n, dtype=torch.uint8),
                      torch.empty(n, dtype=torch.uint8))
    for t in range(steps):
        b.step_host(acts[t], ro, rr, rd, rt, autoreset=True)
        _, o, w, d, tr = seen[t]
        assert torch.equal(o, ro) and torch.equal(w, rr) and torch.equal(d, rd) and torch.equal(tr, rt), t
    assert np.array_equal(a.get_state(), b.get_state())
    assert any(bool(s[4].any()) for s in seen)          # the horizon of 4 was hit inside the rollout
    if kind != "pendulum":
        # compact transport: uint8 actions in, done / truncated as bits out
        nb = (n + 7) // 8
        bits_d = torch.empty((S, nb), dtype=torch.uint8).pin_memory()
        bits_t = torch.empty((S, nb), dtype=torch.uint8).pin_memory()
        seen_c = []

        def on_step_c(t, slot):
            seen_c.append((obs[slot].clone(), rew[slot].clone(), bits_d[slot].clone(), bits_t[slot].clone()))

        c.rollout_host(acts.to(torch.uint8).pin_memory(), obs, rew, bits_d, bits_t, autoreset=True, on_step=on_step_c,
                       u8_actions=True, packed_done=True)
        for t in range(steps):
            o, w, bd, bt = seen_c[t]
            assert torch.equal(o, seen[t][1]) and torch.equal(w, seen[t][2])
            assert np.array_equal(np.unpackbits(bd.numpy(), bitorder="little")[:n], seen[t][3].numpy())
            assert np.array_equal(np.unpackbits(bt.numpy(), bitorder="little")[:n], seen[t][4].numpy())
    else:
        with pytest.raises(Exception):
            c.rollout_host(acts, obs, rew, done, trunc, u8_actions=True)
    for e in (a, b, c):
        e.close()


def test_handle_follows_torch_current_stream(torch, g):
    """core.py binds the handle to torch's CURRENT stream at every step / rollout / reset: actions
    produced on a side stream are consumed by a step issued under the same stream context."""
    n = 1 << 18
    a, b = g.CartPoleEnv(num_envs=n), g.CartPoleEnv(num_envs=n)
    a.reset(seed=4)
    b.reset(seed=4)
    side = torch.cuda.Stream()
    gen_a = torch.Generator(device="cuda").manual_seed(9)
    gen_b = torch.Generator(device="cuda").manual_seed(9)
    for t in range(40):
        with torch.cuda.stream(side):
            act = torch.randint(0, 2, (n,), generator=gen_a, device="cuda", dtype=torch.int32)
            big = torch.randn(1 << 22, device="cuda").sin_()        # keeps the side stream busy before the step
            act = act + (big[:n] * 0).to(torch.int32)
            out = a.step(act, autoreset=True)
        ptr = C_void(a)
        assert ptr == side.cuda_stream
        act_b = torch.randint(0, 2, (n,), generator=gen_b, device="cuda", dtype=torch.int32)
        ob = b.step(act_b, autoreset=True)
        assert C_void(b) == torch.cuda.current_stream().cuda_stream
    side.synchronize()
    a.sync()
    b.sync()
    assert np.array_equal(a.get_state(), b.get_state())
    a.close()
    b.close()


def C_void(env):
    import ctypes
    p = ctypes.c_void_p()
    from gym_rs_b200 import _capi
    _capi.check(_capi.load().gymrs_get_stream(env.handle, ctypes.byref(p)))
    v = p.value or 0
    return 0 if v == 1 else v   # cudaStreamLegacy is how stream 0 is passed down


def test_sharding_invariance_at_full_size(torch, g):
    """results are keyed by GLOBAL env id, so one 1M-env handle and two
    512K-env handles (as two GPUs would hold them) produce identical bits."""
    n = N_FULL
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = [torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(40)]
    whole = g.CartPoleEnv(num_envs=n)
    lo = g.CartPoleEnv(num_envs=n // 2, global_env_offset=0)
    hi = g.CartPoleEnv(num_envs=n // 2, global_env_offset=n // 2)
    for e in (whole, lo, hi):
        e.reset(seed=123)
    for a in acts:
        ow = whole.step(a, autoreset=True)
        ol = lo.step(a[: n // 2], autoreset=True)
        oh = hi.step(a[n // 2:], autoreset=True)
    for e in (whole, lo, hi):
        e.sync()
    assert torch.equal(ow.observation[:, : n // 2], ol.observation)
    assert torch.equal(ow.observation[:, n // 2:], oh.observation)
    assert torch.equal(ow.done[: n // 2], ol.done) and torch.equal(ow.done[n // 2:], oh.done)
    # checksum of checksums: the whole-batch sum equals the sum of the shard sums exactly (u8 counts)
    assert int(ow.done.sum()) == int(ol.done.sum()) + int(oh.done.sum())
    for e in (whole, lo, hi):
        e.close()


def test_mirror_symmetry_at_full_size(torch, g):
    """Size-independent property at BASELINE's full size: CartPole's dynamics are odd under
    (state, force) -> (-state, -force).  Every operation on the device path (polynomial sin/cos,
    FMA chain, reciprocal + Newton step, |x| > T tests) is sign-symmetric, so stepping the mirrored
    batch must give the exactly negated state and identical done flags -- for all 2^20 envs, over
    several steps, with no oracle involved."""
    n = N_FULL
    st, act = cartpole_inputs(n, seed=13)
    a = g.CartPoleEnv(num_envs=n)
    b = g.CartPoleEnv(num_envs=n)
    a.set_state(st)
    b.set_state(-st)
    gen = torch.Generator(device="cuda").manual_seed(9)
    for _ in range(12):
        acts = torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int32)
        oa = a.step(acts)
        ob = b.step(1 - acts)
    a.sync()
    b.sync()
    assert torch.equal(oa.observation, -ob.observation)
    assert torch.equal(oa.done, ob.done) and torch.equal(oa.reward, ob.reward)
    assert torch.equal(a.steps_beyond_terminated, b.steps_beyond_terminated)
    a.close()
    b.close()
    # MountainCar: outputs stay inside the observation space for arbitrary inputs (clip, mountain_car.rs:413-416)
    m = g.MountainCarEnv(num_envs=n)
    wild = np.stack([np.random.default_rng(1).uniform(-5, 5, n), np.random.default_rng(2).uniform(-1, 1, n)])
    m.set_state(wild.astype(np.float32))
    out = m.step(torch.randint(0, 3, (n,), generator=gen, device="cuda", dtype=torch.int32))
    m.sync()
    o = out.observation
    assert float(o[0].min()) >= np.float32(-1.2) and float(o[0].max()) <= np.float32(0.6)
    assert float(o[1].abs().max()) <= np.float32(0.07)
    assert bool(((o[0] > np.float32(-1.2)) | (o[1] >= 0)).all())        # wall rule
    assert torch.equal(out.done.bool(), (o[0] >= 0.5) & (o[1] >= 0))
    m.close()


def test_determinism_same_seed_same_bits(torch, g):
    n = 1 << 18
    gen = torch.Generator(device="cuda").manual_seed(4)
    acts = [torch.randint(0, 3, (n,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(25)]
    outs = []
    for _ in range(2):
        env = g.MountainCarEnv(num_envs=n)
        env.reset(seed=99)
        for a in acts:
            out = env.step(a, autoreset=True)
        env.sync()
        outs.append(out.observation.clone())
        env.close()
    assert torch.equal(outs[0], outs[1])


# --------------------------------------------------------------------------------------
# errors, time limit, clone, properties, scalar Env surface
# --------------------------------------------------------------------------------------

def test_invalid_action_is_reported_and_env_left_untouched(torch, g):
    n = 4096
    st, act = cartpole_inputs(n, seed=12)
    act[1234] = 2    # not in Discrete(2)
    act[99] = -1     # not a usize
    env = g.CartPoleEnv(num_envs=n, global_env_offset=500)
    env.set_state(st)
    out = env.step(dev_actions(torch, act))
    # the reference's panic text "{} usize invalid" (cartpole.rs:404), plus the env it happened in
    # (two offenders race for the report slot: either may win)
    with pytest.raises(AssertionError, match=r"^(2|-1) usize invalid \(env (1734|599)\)$"):
        env.sync()
    env.sync()  # sticky flag is cleared once reported
    got = out.observation.cpu().numpy()
    assert np.array_equal(got[:, 1234], st[:, 1234]) and np.array_equal(got[:, 99], st[:, 99])
    ok = np.ones(n, dtype=bool)
    ok[[99, 1234]] = False
    a2 = act.copy()
    a2[~ok] = 0
    ref = oracle.step_batch(oracle.CARTPOLE, st, a2)
    assert_within(got[:, ok], ref["state"][:, ok], "valid envs still stepped")
    env.close()
    mc = g.MountainCarEnv(num_envs=8)
    mc.step(dev_actions(torch, np.array([0, 1, 2, 3, 0, 1, 2, 0], dtype=np.int32)))
    with pytest.raises(AssertionError):
        mc.sync()
    mc.close()


def test_time_limit_truncation(torch, g):
    n = 2048
    env = g.MountainCarEnv(num_envs=n, time_limit=True)
    p = env.params
    p.max_episode_steps = 7
    env.params = p
    env.reset(seed=1)
    a = dev_actions(torch, np.ones(n, dtype=np.int32))
    for t in range(1, 8):
        out = env.step(a, autoreset=True)
        env.sync()
        tr = out.truncated.cpu().numpy()
        assert tr.all() == (t == 7) and tr.any() == (t == 7)
    st = env.get_state()
    assert st[0].min() >= -0.6 and st[0].max() < -0.4 and np.all(st[1] == 0)  # truncated envs were reset
    out = env.step(a, autoreset=True)
    env.sync()
    assert not out.truncated.cpu().numpy().any()
    env.close()
    # default (reference behaviour): never truncated (cartpole.rs:480, mountain_car.rs:432)
    env = g.MountainCarEnv(num_envs=n)
    for _ in range(250):
        out = env.step(a)
    env.sync()
    assert not out.truncated.cpu().numpy().any()
    env.close()


def test_pendulum_host_step_rollout_and_time_limit(torch, g):
    """Continuous-action env through every entry point: host step (float actions), fused rollout with
    TIME_LIMIT truncation + auto-reset, and theta staying wrapped over a long spin."""
    n = 300001
    r = np.random.default_rng(4)
    env = g.PendulumEnv(num_envs=n, time_limit=True)
    p = env.params
    p.max_episode_steps = 9
    env.params = p
    env.reset(seed=6)
    st0 = env.get_state()
    act = r.uniform(-2.5, 2.5, n).astype(np.float32)
    obs = np.empty((3, n), dtype=np.float32)
    rew = np.empty(n, dtype=np.float32)
    done = np.empty(n, dtype=np.uint8)
    trunc = np.empty(n, dtype=np.uint8)
    env.step_host(act, obs, rew, done, trunc, autoreset=True)
    ref = oracle.step_batch(oracle.PENDULUM, st0, act)
    assert_within(obs, ref["obs"], "pendulum host step obs")
    assert_within(rew, ref["reward"], "pendulum host step reward")
    assert not done.any() and not trunc.any()
    # 20 fused steps with a horizon of 9: truncation at elapsed == 9 and 18 (1 host step + k rollout steps)
    k = 20
    acts = torch.as_tensor(r.uniform(-2, 2, (k, n)).astype(np.float32)).cuda()
    done_out = torch.zeros((k, n), device="cuda", dtype=torch.uint8)
    obs_out = torch.zeros((k, 3, n), device="cuda")
    env.rollout(acts, obs_out, None, done_out, autoreset=True)
    env.sync()
    assert not done_out.any()                       # Pendulum never terminates
    el = core_elapsed(env)
    assert (el == (1 + k) % 9).all()                # 21 steps, reset at 9 and 18 -> 3 steps into the third episode
    o = obs_out[-1].cpu().numpy()
    assert np.abs(o[0] ** 2 + o[1] ** 2 - 1).max() < 1e-5 and np.abs(o[2]).max() <= 8.0
    # constant maximal torque spins the pendulum up; the stored angle must stay wrapped
    spin = g.PendulumEnv(num_envs=4096)
    spin.reset(seed=1)
    full = torch.full((4096,), 2.0, device="cuda")
    for _ in range(600):
        spin.step(full)
    spin.sync()
    th = spin.get_state()[0]
    assert np.abs(th).max() <= math.pi + 1e-5 and np.isfinite(th).all()
    env.close()
    spin.close()


def core_elapsed(env):
    """device view of the TIME_LIMIT step counters (gymrs_buffers.elapsed_steps)"""
    from gym_rs_b200.core import _as_tensor
    return _as_tensor(env._buf.elapsed_steps, (env.num_envs,), "<i4", None, env.device).to("cpu").numpy().astype(np.int64)


def test_clone_and_stream_switch_keep_chained_state_consistent(torch, g):
    """Clone after chained (pdl = 2) steps, then keep stepping both; switch the original to another
    stream in the middle.  Both must track a plainly-launched reference handle bit for bit."""
    n = 1 << 18
    gen = torch.Generator(device="cuda").manual_seed(2)
    acts = [torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(6)]
    torch.cuda.synchronize()
    a = g.CartPoleEnv(num_envs=n)
    ref = g.CartPoleEnv(num_envs=n)
    a.set_launch_config(pdl=2)
    ref.set_launch_config(pdl=0)
    for e in (a, ref):
        e.reset(seed=17)
    for t in range(30):
        a.step(acts[t % 6], autoreset=True)
        ref.step(acts[t % 6], autoreset=True)
    twin = a.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    a.set_stream(side.cuda_stream)
    for t in range(30, 60):
        with torch.cuda.stream(side):
            a.step(acts[t % 6], autoreset=True)
        twin.step(acts[t % 6], autoreset=True)
        ref.step(acts[t % 6], autoreset=True)
    for e in (a, twin, ref):
        e.sync()
    s = ref.get_state()
    assert np.array_equal(a.get_state(), s) and np.array_equal(twin.get_state(), s)
    for e in (a, twin, ref):
        e.close()


def test_clone_is_a_deep_copy(torch, g):
    n = 10000
    env = g.CartPoleEnv(num_envs=n)
    env.reset(seed=5)
    a = dev_actions(torch, np.ones(n, dtype=np.int32))
    env.step(a)
    twin = env.clone()
    s0 = env.get_state()
    assert np.array_equal(twin.get_state(), s0)
    env.step(a)
    assert np.array_equal(twin.get_state(), s0)
    o1 = twin.step(a).observation.cpu().numpy()
    env.sync()
    assert np.array_equal(o1, env.get_state())
    d = env.serialize()
    assert d["num_envs"] == n and len(d["state"]) == 4 and d["params"]["gravity"] == 9.8
    env.close()
    twin.close()


def test_env_properties(torch, g):
    cp = g.CartPoleEnv()
    assert cp.action_space() == g.Discrete(2)                      # cartpole.rs:112
    hi = cp.observation_space().high
    assert hi.to_vec() == [4.8, math.inf, 0.41887902047863906, math.inf]   # cartpole.rs:105-113
    assert cp.observation_space().low == -hi
    assert cp.reward_range() == g.RewardRange(-math.inf, math.inf)  # core.rs:16-19
    assert cp.render_mode() == g.RenderMode.NONE and cp.metadata().render_fps == 50
    cp.close()
    mc = g.MountainCarEnv()
    assert mc.action_space() == g.Discrete(3)                      # mountain_car.rs:363
    sp = mc.observation_space()
    assert sp.low.to_vec() == [-1.2, -0.07] and sp.high.to_vec() == [0.6, 0.07]  # :353-354
    assert mc.metadata().render_fps == 30
    mc.close()
    pd = g.PendulumEnv()
    assert pd.action_space() == g.BoxR(-2.0, 2.0)
    assert pd.observation_space().high.to_vec() == [1.0, 1.0, 8.0]
    pd.close()
    with pytest.raises(ValueError):
        g.CartPoleEnv(g.RenderMode.Human)


def test_scalar_env_surface_like_examples_cartpole(torch, g):
    """BASELINE config 1 / examples/cartpole.rs:7-33 with RenderMode::None: one env, random actions,
    15 episodes of <= 475 steps, reset after each; cross-checked step by step with the oracle."""
    import random
    rng = random.Random(0)
    env = g.CartPoleEnv(g.RenderMode.NONE)
    env.reset(None, False, None)
    rewards = []
    total_steps = 0
    for ep in range(15):
        state, _ = env.reset(seed=ep)
        o = oracle.CartPoleEnv()
        oracle.lib().orc_cartpole_new(oracle.C.byref(o))
        for k, v in enumerate(state.to_vec()):
            o.state[k] = v
        current_reward = 0.0
        for _ in range(475):
            action = rng.randrange(2)
            sr = env.step(action)
            rew, dn, tr = oracle.C.c_double(), oracle.C.c_int(), oracle.C.c_int()
            oracle.lib().orc_cartpole_step(oracle.C.byref(o), action, oracle.C.byref(rew),
                                           oracle.C.byref(dn), oracle.C.byref(tr))
            assert_within(sr.observation.to_vec(), list(o.state), "scalar step")
            for k, v in enumerate(sr.observation.to_vec()):
                o.state[k] = v  # resync
            assert sr.reward == rew.value and sr.truncated is False and sr.info == ()
            current_reward += sr.reward
            total_steps += 1
            if sr.done:
                assert env.steps_beyond_terminated == 0
                break
        rewards.append(current_reward)
    assert len(rewards) == 15 and all(5 <= r <= 475 for r in rewards)
    with pytest.raises(AssertionError, match="2 usize invalid"):   # cartpole.rs:402-406
        env.step(2)
    env.close()
    mc = g.MountainCarEnv(g.RenderMode.NONE)
    mc.reset(seed=0)
    sr = mc.step(1)
    assert sr.reward == -1.0 and sr.done is False and sr.info is None
    with pytest.raises(AssertionError, match=r"3 \(usize\) invalid"):  # mountain_car.rs:402-406
        mc.step(3)
    mc.close()


def test_stepping_after_termination_warns_like_the_reference(torch, g, caplog):
    """cartpole.rs:455-464: the first terminal step still pays 1.0; every later step pays 0.0, bumps
    steps_beyond_terminated and logs a warning (log::warn!, :461).  MountainCar has no such state."""
    import logging
    env = g.CartPoleEnv(g.RenderMode.NONE)
    env.reset(seed=3)
    with caplog.at_level(logging.WARNING, logger="gym_rs"):
        for _ in range(400):
            sr = env.step(1)
            if sr.done:
                break
        assert sr.done and sr.reward == 1.0 and env.steps_beyond_terminated == 0
        assert not caplog.records
        sr = env.step(1)
        assert sr.done and sr.reward == 0.0 and env.steps_beyond_terminated == 1
        assert len(caplog.records) == 1 and "after termination" in caplog.records[0].getMessage()
        env.reset(seed=4)
        env.step(0)
        assert len(caplog.records) == 1
    env.close()


def test_num_envs_one_reset_returns_the_observation_type(torch, g):
    """core.rs:45-50: reset returns (Observation, Option<ResetInfo>) -- for Pendulum the observation is
    (cos, sin, theta_dot), not the (theta, theta_dot) state."""
    from gym_rs_b200.envs.classical_control.pendulum import PendulumObservation
    env = g.PendulumEnv()
    obs, info = env.reset(seed=5, return_info=True)
    assert isinstance(obs, PendulumObservation) and info == ()
    st = env.get_state()[:, 0]
    assert abs(obs.to_vec()[0] - math.cos(st[0])) < 1e-6 and abs(obs.to_vec()[1] - math.sin(st[0])) < 1e-6
    assert abs(obs.to_vec()[2] - st[1]) < 1e-7
    sr = env.step(0.5)
    assert isinstance(sr.observation, PendulumObservation)
    env.close()
