#!/usr/bin/env python
"""Generate tests/golden/*.json -- known-answer vectors for the oracle.

The reference (a Rust crate) cannot be run in this image and has no test that
calls step/reset, so these vectors are produced by an INDEPENDENT evaluation of
the formulas cited from the reference source:

  * ``mp``  : 50-digit mpmath evaluation (no f64 rounding at all), and
  * ``py``  : a straight Python-float (IEEE f64, libm sin/cos) evaluation in the
              reference's operation order,

neither of which shares code with oracle/gymrs_oracle.c.  The committed JSON
holds the ``py`` value (what the Rust f64 code computes, up to libm sin/cos
last-bit differences) plus the ``mp`` value; tests/test_oracle.py requires the
C oracle to equal ``py`` to <= 2 ulp and ``mp`` to 1e-13 relative.

Formulas (file:line in /root/reference/src/envs/classical_control/):
  CartPole     cartpole.rs:414-453   (PML = masspole + length, :150-152)
  MountainCar  mountain_car.rs:411-423, clip = utils/custom/util_fns.rs:2-10
  Pendulum     upstream Gym pendulum.py -- not in reference

Run:  python tests/golden/make_golden.py      (needs mpmath; CPU only)
"""
import json
import math
import os
import random

import mpmath as mp

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))

# ---- the inputs listed in SURVEY.md Appendix B, plus seeded random ones -----
CARTPOLE_FIXED = [
    ((0.0, 0.0, 0.0, 0.0), 1),
    ((0.0, 0.0, 0.0, 0.0), 0),
    ((0.01, -0.02, 0.03, 0.04), 1),
    ((0.01, -0.02, 0.03, 0.04), 0),
    ((2.39, 1.0, 0.0, 0.0), 1),
    ((0.0, 0.0, 0.2, 1.5), 0),
    ((-1.0, -2.0, -0.15, -1.0), 1),
    ((-2.39, -1.0, 0.0, 0.0), 0),
    ((0.0, 0.0, -0.2, -1.5), 1),
    ((2.4, 0.0, 0.0, 0.0), 1),       # x == threshold exactly: strict '>' -> not done
]
MOUNTAIN_CAR_FIXED = [
    ((-0.5, 0.0), 0), ((-0.5, 0.0), 1), ((-0.5, 0.0), 2),
    ((-1.2, -0.05), 0), ((-1.19, -0.07), 0),
    ((0.49, 0.07), 2), ((0.6, 0.07), 2), ((0.45, 0.04), 1),
    ((0.5, 0.0), 1),                  # goal position reached with v that turns negative
    ((-1.2, 0.0), 2),
]
PENDULUM_FIXED = [
    ((0.0, 0.0), 0.0), ((math.pi, 0.0), 0.0), ((1.0, 1.0), 2.0), ((1.0, 1.0), 5.0),
    ((-3.0, 7.9), 2.0), ((3.0, -7.9), -2.0), ((10.0, 0.5), -0.3), ((-10.0, -0.5), 0.3),
]


def cartpole_py(s, a, semi_implicit=False):
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    thr_th = 12. * 2. * math.pi / 360.
    thr_x = 2.4
    x, x_dot, theta, theta_dot = s
    force = force_mag if a == 1 else -force_mag
    c, sn = math.cos(theta), math.sin(theta)
    M = masspole + masscart
    pml = masspole + length
    temp = (force + pml * (theta_dot * theta_dot) * sn) / M
    thetaacc = (gravity * sn - c * temp) / (length * (4.0 / 3.0 - masspole * (c * c) / M))
    xacc = temp - pml * thetaacc * c / M
    if not semi_implicit:
        x = x + tau * x_dot
        x_dot = x_dot + tau * xacc
        theta = theta + tau * theta_dot
        theta_dot = theta_dot + tau * thetaacc
    else:
        x_dot = x_dot + tau * xacc
        x = x + tau * x_dot
        theta_dot = theta_dot + tau * thetaacc
        theta = theta + tau * theta_dot
    done = x < -thr_x or x > thr_x or theta < -thr_th or theta > thr_th
    return [x, x_dot, theta, theta_dot], bool(done)


def cartpole_mp(s, a, semi_implicit=False):
    f = mp.mpf
    gravity, masscart, masspole, length, force_mag, tau = map(f, (9.8, 1.0, 0.1, 0.5, 10.0, 0.02))
    x, x_dot, theta, theta_dot = map(f, s)
    force = force_mag if a == 1 else -force_mag
    c, sn = mp.cos(theta), mp.sin(theta)
    M = masspole + masscart
    pml = masspole + length
    temp = (force + pml * theta_dot ** 2 * sn) / M
    thetaacc = (gravity * sn - c * temp) / (length * (f(4.0 / 3.0) - masspole * c ** 2 / M))
    xacc = temp - pml * thetaacc * c / M
    if not semi_implicit:
        x, x_dot, theta, theta_dot = (x + tau * x_dot, x_dot + tau * xacc,
                                      theta + tau * theta_dot, theta_dot + tau * thetaacc)
    else:
        x_dot = x_dot + tau * xacc
        x = x + tau * x_dot
        theta_dot = theta_dot + tau * thetaacc
        theta = theta + tau * theta_dot
    return [float(v) for v in (x, x_dot, theta, theta_dot)]


def clip(v, lo, hi):
    if lo <= v <= hi:
        return v
    elif v > hi:
        return hi
    return lo


def mountain_car_py(s, a):
    p, v = s
    v = v + ((float(a) - 1.) * 0.001 + math.cos(3. * p) * (-0.0025))
    v = clip(v, -0.07, 0.07)
    p = p + v
    p = clip(p, -1.2, 0.6)
    if p == -1.2 and v < 0.:
        v = 0.
    done = p >= 0.5 and v >= 0.
    return [p, v], bool(done)


def mountain_car_mp(s, a):
    f = mp.mpf
    p, v = map(f, s)
    v = v + ((f(a) - 1) * f(0.001) + mp.cos(3 * p) * (-f(0.0025)))
    v = clip(v, f(-0.07), f(0.07))
    p = p + v
    p = clip(p, f(-1.2), f(0.6))
    if p == f(-1.2) and v < 0:
        v = f(0)
    return [float(p), float(v)]


def pendulum_py(s, u):
    max_speed, max_torque, dt, g, m, l = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0
    th, thdot = s
    u = clip(u, -max_torque, max_torque)
    y = th + math.pi
    an = (y - 2. * math.pi * math.floor(y / (2. * math.pi))) - math.pi
    costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u)
    newthdot = thdot + (3. * g / (2. * l) * math.sin(th) + 3.0 / (m * (l * l)) * u) * dt
    newthdot = clip(newthdot, -max_speed, max_speed)
    newth = th + newthdot * dt
    return [newth, newthdot], [math.cos(newth), math.sin(newth), newthdot], -costs


def pendulum_mp(s, u):
    f = mp.mpf
    max_speed, max_torque, dt, g, m, l = map(f, (8.0, 2.0, 0.05, 10.0, 1.0, 1.0))
    th, thdot = map(f, s)
    u = clip(f(u), -max_torque, max_torque)
    pi = f(math.pi)  # the f64 constant, as the program would use
    y = th + pi
    an = (y - 2 * pi * mp.floor(y / (2 * pi))) - pi
    costs = an ** 2 + f(0.1) * thdot ** 2 + f(0.001) * u ** 2
    newthdot = thdot + (3 * g / (2 * l) * mp.sin(th) + 3 / (m * l ** 2) * u) * dt
    newthdot = clip(newthdot, -max_speed, max_speed)
    newth = th + newthdot * dt
    return ([float(newth), float(newthdot)],
            [float(mp.cos(newth)), float(mp.sin(newth)), float(newthdot)], float(-costs))


def main():
    rnd = random.Random(20261017)
    out = {"cartpole": [], "cartpole_semi_implicit": [], "mountain_car": [], "pendulum": []}

    cp_inputs = list(CARTPOLE_FIXED)
    for _ in range(40):
        cp_inputs.append(((rnd.uniform(-2.4, 2.4), rnd.uniform(-3, 3), rnd.uniform(-0.21, 0.21),
                           rnd.uniform(-3, 3)), rnd.randrange(2)))
    for s, a in cp_inputs:
        st, done = cartpole_py(s, a)
        out["cartpole"].append(dict(state=list(s), action=a, next_state=st, done=done,
                                    next_state_mp=cartpole_mp(s, a)))
        st, done = cartpole_py(s, a, True)
        out["cartpole_semi_implicit"].append(dict(state=list(s), action=a, next_state=st, done=done,
                                                  next_state_mp=cartpole_mp(s, a, True)))

    mc_inputs = list(MOUNTAIN_CAR_FIXED)
    for _ in range(40):
        mc_inputs.append(((rnd.uniform(-1.2, 0.6), rnd.uniform(-0.07, 0.07)), rnd.randrange(3)))
    for s, a in mc_inputs:
        st, done = mountain_car_py(s, a)
        out["mountain_car"].append(dict(state=list(s), action=a, next_state=st, done=done,
                                        next_state_mp=mountain_car_mp(s, a)))

    pd_inputs = list(PENDULUM_FIXED)
    for _ in range(40):
        pd_inputs.append(((rnd.uniform(-math.pi, math.pi), rnd.uniform(-8, 8)), rnd.uniform(-2.5, 2.5)))
    for s, u in pd_inputs:
        st, obs, rew = pendulum_py(s, u)
        st_mp, obs_mp, rew_mp = pendulum_mp(s, u)
        out["pendulum"].append(dict(state=list(s), action=u, next_state=st, obs=obs, reward=rew,
                                    next_state_mp=st_mp, obs_mp=obs_mp, reward_mp=rew_mp))

    # CartPole reward sequence past termination (cartpole.rs:455-464): fixed action 1
    # from rest drives x past 2.4; rewards 1.0 ... 1.0(first done) then 0.0.
    s = (0.0, 0.0, 0.0, 0.0)
    seq = []
    sbt = None
    for _ in range(60):
        s, done = cartpole_py(s, 1)
        if not done:
            r = 1.0
        elif sbt is None:
            sbt = 0
            r = 1.0
        else:
            sbt += 1
            r = 0.0
        seq.append(dict(state=list(s), done=done, reward=r))
    out["cartpole_reward_sequence"] = seq

    # Philox4x32-10 known-answer vectors (Random123 kat_vectors)
    out["philox4x32_10"] = [
        dict(ctr=[0, 0, 0, 0], key=[0, 0],
             out=[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        dict(ctr=[0xffffffff] * 4, key=[0xffffffff] * 2,
             out=[0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        dict(ctr=[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], key=[0xa4093822, 0x299f31d0],
             out=[0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]

    with open(os.path.join(HERE, "step_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "step_vectors.json"),
          {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
