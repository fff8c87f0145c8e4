"""A second, independent restatement of the reference's step arithmetic -- plain Python floats (IEEE f64, the same
libm sin / cos as the C oracle), written straight from the reference source, NOT from oracle/gymrs_oracle.c --
checked bit for bit against the C oracle on random inputs.

The reference cannot be compiled in this image (no cargo), so no output of the crate itself pins the oracle
(DESIGN.md section 5).  What this test adds is that two restatements in two languages, written separately from the
same source lines, agree to the last bit on tens of thousands of random (state, action) pairs: a transcription slip
in either one would have to be made twice to survive.

CPU only; pure-Python loops, sized to run in about a second.
"""
import math

import numpy as np
import pytest

import oracle

N = 20000


# ---- reference: src/envs/classical_control/cartpole.rs ---------------------------------------------------
def cartpole_step(x, x_dot, theta, theta_dot, action, semi_implicit=False):
    # constants: cartpole.rs:94-103; helpers total_mass :146-148, polemass_length :150-152 (a SUM, sic)
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    total_mass = masspole + masscart
    polemass_length = masspole + length
    theta_threshold_radians = 12.0 * 2.0 * math.pi / 360.0
    x_threshold = 2.4
    force = force_mag if action == 1 else -force_mag                                           # :414-418
    costheta = math.cos(theta)                                                                 # :420
    sintheta = math.sin(theta)                                                                 # :421
    # powf(2.) of an f64 is x * x (LLVM folds pow(x, 2.0) to a multiply without any fast-math flag)
    temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass         # :423-424
    thetaacc = (gravity * sintheta - costheta * temp) / (
        length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass))                  # :425-428
    xacc = temp - polemass_length * thetaacc * costheta / total_mass                           # :429
    if not semi_implicit:                                                                      # :431-436
        x = x + tau * x_dot
        x_dot = x_dot + tau * xacc
        theta = theta + tau * theta_dot
        theta_dot = theta_dot + tau * thetaacc
    else:                                                                                      # :437-441
        x_dot = x_dot + tau * xacc
        x = x + tau * x_dot
        theta_dot = theta_dot + tau * thetaacc
        theta = theta + tau * theta_dot
    done = (x < -x_threshold or x > x_threshold or theta < -theta_threshold_radians
            or theta > theta_threshold_radians)                                                # :450-453
    return (x, x_dot, theta, theta_dot), 1.0, done   # reward of a live or first-terminal step :455-459


# ---- reference: src/utils/custom/util_fns.rs:2-10 --------------------------------------------------------
def clip(value, left_bound, right_bound):
    if left_bound <= value <= right_bound:
        return value
    if value > right_bound:
        return right_bound
    return left_bound


# ---- reference: src/envs/classical_control/mountain_car.rs -----------------------------------------------
def mountain_car_step(position, velocity, action):
    # constants: mountain_car.rs:344-351
    min_position, max_position, max_speed, goal_position, goal_velocity = -1.2, 0.6, 0.07, 0.5, 0.0
    force, gravity = 0.001, 0.0025
    velocity += (float(action) - 1.0) * force + math.cos(3.0 * position) * (-gravity)          # :411-412
    velocity = clip(velocity, -max_speed, max_speed)                                           # :413
    position += velocity                                                                       # :415
    position = clip(position, min_position, max_position)                                      # :416
    if position == min_position and velocity < 0.0:                                            # :418-420
        velocity = 0.0
    done = position >= goal_position and velocity >= goal_velocity                             # :422
    return (position, velocity), -1.0, done                                                    # :423


# ---- Pendulum-v1: SURVEY.md Appendix D (upstream Gym pendulum.py; not in the reference) --------------------
def pendulum_step(th, thdot, u):
    max_speed, max_torque, dt, g, m, l = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0
    u = min(max(u, -max_torque), max_torque)
    norm = ((th + math.pi) % (2.0 * math.pi)) - math.pi      # floor-mod for a positive modulus, like numpy's
    cost = norm * norm + 0.1 * (thdot * thdot) + 0.001 * (u * u)
    newthdot = thdot + (3.0 * g / (2.0 * l) * math.sin(th) + 3.0 / (m * l * l) * u) * dt
    newthdot = min(max(newthdot, -max_speed), max_speed)
    newth = th + newthdot * dt
    return (newth, newthdot), (math.cos(newth), math.sin(newth), newthdot), -cost


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("semi", [False, True])
def test_cartpole_second_restatement_is_bit_identical(semi):
    r = np.random.default_rng(41)
    st = np.stack([r.uniform(-2.6, 2.6, N), r.uniform(-3, 3, N), r.uniform(-0.25, 0.25, N), r.uniform(-3, 3, N)])
    st[2, ::50] = r.uniform(-40.0, 40.0, len(st[2, ::50]))  # poles far beyond the threshold: full-range sin / cos
    act = r.integers(0, 2, N).astype(np.int32)
    p = oracle.default_params(oracle.CARTPOLE)
    p.kinematics_integrator = 1 if semi else 0
    ref = oracle.step_batch(oracle.CARTPOLE, st, act, params=p)
    mine = [cartpole_step(*st[:, i], int(act[i]), semi) for i in range(N)]
    assert np.array_equal(_bits([m[0] for m in mine]).T, _bits(ref["state"]))
    assert np.array_equal(np.array([m[2] for m in mine], dtype=np.uint8), ref["done"])
    assert np.all(ref["reward"] == 1.0)
    assert 0 < ref["done"].sum() < N


def test_mountain_car_second_restatement_is_bit_identical():
    r = np.random.default_rng(42)
    st = np.stack([r.uniform(-1.2, 0.6, N), r.uniform(-0.07, 0.07, N)])
    st[0, ::40] = -1.2   # at the wall
    st[0, 1::40] = 0.6   # at the right end
    st[0, 2::40] = r.uniform(0.45, 0.6, len(st[0, 2::40]))  # around the goal
    act = r.integers(0, 3, N).astype(np.int32)
    ref = oracle.step_batch(oracle.MOUNTAIN_CAR, st, act)
    mine = [mountain_car_step(st[0, i], st[1, i], int(act[i])) for i in range(N)]
    assert np.array_equal(_bits([m[0] for m in mine]).T, _bits(ref["state"]))
    assert np.array_equal(np.array([m[2] for m in mine], dtype=np.uint8), ref["done"])
    assert np.all(ref["reward"] == -1.0)
    assert 0 < ref["done"].sum() < N


def test_pendulum_second_restatement_is_bit_identical():
    r = np.random.default_rng(43)
    st = np.stack([r.uniform(-math.pi, math.pi, N), r.uniform(-8, 8, N)])
    st[0, ::25] = r.uniform(-30.0, 30.0, len(st[0, ::25]))  # unwrapped angles: angle_normalize does the work
    act = r.uniform(-2.5, 2.5, N)
    ref = oracle.step_batch(oracle.PENDULUM, st, act)
    mine = [pendulum_step(st[0, i], st[1, i], float(act[i])) for i in range(N)]
    assert np.array_equal(_bits([m[0] for m in mine]).T, _bits(ref["state"]))
    assert np.array_equal(_bits([m[1] for m in mine]).T, _bits(ref["obs"]))
    assert np.array_equal(_bits([m[2] for m in mine]), _bits(ref["reward"]))
    assert not ref["done"].any()
