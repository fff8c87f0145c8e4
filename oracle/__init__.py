"""ctypes binding of the CPU oracle (libgymrs_oracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package
``gym_rs_b200`` never imports this module (tests/test_boundary.py checks that).

See oracle/gymrs_oracle.h for what is restated and the parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgymrs_oracle.so")

CARTPOLE, MOUNTAIN_CAR, PENDULUM = 0, 1, 2
STATE_DIM = {CARTPOLE: 4, MOUNTAIN_CAR: 2, PENDULUM: 2}
OBS_DIM = {CARTPOLE: 4, MOUNTAIN_CAR: 2, PENDULUM: 3}


class CartPoleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "gravity", "masscart", "masspole", "length", "force_mag", "tau",
        "theta_threshold_radians", "x_threshold")] + [("kinematics_integrator", C.c_int)]


class MountainCarParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "min_position", "max_position", "max_speed", "goal_position",
        "goal_velocity", "force", "gravity")]


class PendulumParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("max_speed", "max_torque", "dt", "g", "m", "l")]


class CartPoleEnv(C.Structure):
    _fields_ = [("p", CartPoleParams), ("state", C.c_double * 4),
                ("steps_beyond_terminated", C.c_longlong)]


class MountainCarEnv(C.Structure):
    _fields_ = [("p", MountainCarParams), ("state", C.c_double * 2)]


class PendulumEnv(C.Structure):
    _fields_ = [("p", PendulumParams), ("state", C.c_double * 2)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("gymrs_oracle.c", "gymrs_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libgymrs_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    dp = C.POINTER(C.c_double)
    L.orc_clip.restype = C.c_double
    L.orc_clip.argtypes = [C.c_double] * 3
    L.orc_clip_i64.restype = C.c_longlong
    L.orc_clip_i64.argtypes = [C.c_longlong] * 3
    L.orc_discrete_contains.restype = C.c_int
    L.orc_discrete_contains.argtypes = [C.c_size_t, C.c_size_t]
    L.orc_rand_random.restype = C.c_uint64
    L.orc_rand_random.argtypes = [C.c_int, C.c_uint64]
    L.orc_philox4x32_10.restype = None
    L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_uniform_from_word.restype = C.c_double
    L.orc_uniform_from_word.argtypes = [C.c_uint32, C.c_double, C.c_double]
    L.orc_angle_normalize.restype = C.c_double
    L.orc_angle_normalize.argtypes = [C.c_double]

    L.orc_cartpole_default_params.argtypes = [C.POINTER(CartPoleParams)]
    L.orc_cartpole_new.argtypes = [C.POINTER(CartPoleEnv)]
    L.orc_cartpole_step.restype = C.c_int
    L.orc_cartpole_step.argtypes = [C.POINTER(CartPoleEnv), C.c_size_t, dp,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_cartpole_reset.argtypes = [C.POINTER(CartPoleEnv), C.c_uint64, C.c_uint64,
                                     C.c_uint64, dp, dp]
    L.orc_cartpole_observation_space.argtypes = [C.POINTER(CartPoleParams), dp, dp]

    L.orc_mountain_car_default_params.argtypes = [C.POINTER(MountainCarParams)]
    L.orc_mountain_car_new.argtypes = [C.POINTER(MountainCarEnv)]
    L.orc_mountain_car_step.restype = C.c_int
    L.orc_mountain_car_step.argtypes = [C.POINTER(MountainCarEnv), C.c_size_t, dp,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_mountain_car_reset.argtypes = [C.POINTER(MountainCarEnv), C.c_uint64, C.c_uint64,
                                         C.c_uint64, dp, dp]
    L.orc_mountain_car_observation_space.argtypes = [C.POINTER(MountainCarParams), dp, dp]

    L.orc_pendulum_default_params.argtypes = [C.POINTER(PendulumParams)]
    L.orc_pendulum_new.argtypes = [C.POINTER(PendulumEnv)]
    L.orc_pendulum_step.restype = C.c_int
    L.orc_pendulum_step.argtypes = [C.POINTER(PendulumEnv), C.c_double, dp, dp,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_pendulum_reset.argtypes = [C.POINTER(PendulumEnv), C.c_uint64, C.c_uint64,
                                     C.c_uint64, dp, dp]
    L.orc_pendulum_observation_space.argtypes = [C.POINTER(PendulumParams), dp, dp]

    L.orc_step_batch.restype = C.c_longlong
    L.orc_step_batch.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_reset_batch.restype = None
    L.orc_reset_batch.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_uint64, C.c_uint64,
                                  C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_bench_rollout.restype = C.c_double
    L.orc_bench_rollout.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_uint64, dp]
    L.orc_bench_regions.restype = C.c_double
    L.orc_bench_regions.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_uint64, dp, dp]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(kind: int):
    L = lib()
    if kind == CARTPOLE:
        p = CartPoleParams()
        L.orc_cartpole_default_params(C.byref(p))
    elif kind == MOUNTAIN_CAR:
        p = MountainCarParams()
        L.orc_mountain_car_default_params(C.byref(p))
    else:
        p = PendulumParams()
        L.orc_pendulum_default_params(C.byref(p))
    return p


def step_batch(kind, state, actions, sbt=None, params=None):
    """Step n independent envs.  ``state`` is [state_dim, n] (any float dtype; it is
    widened to f64, so passing the device's f32 state gives both sides identical
    inputs).  Returns dict(state, obs, reward, done, sbt, invalid)."""
    L = lib()
    st = np.ascontiguousarray(np.asarray(state, dtype=np.float64)).copy()
    sd, n = st.shape
    assert sd == STATE_DIM[kind]
    if kind == PENDULUM:
        act = np.ascontiguousarray(np.asarray(actions, dtype=np.float64))
    else:
        act = np.ascontiguousarray(np.asarray(actions, dtype=np.int32))
    assert act.shape == (n,)
    obs = np.zeros((OBS_DIM[kind], n), dtype=np.float64)
    reward = np.zeros(n, dtype=np.float64)
    done = np.zeros(n, dtype=np.uint8)
    sbt_arr = None
    if sbt is not None:
        sbt_arr = np.ascontiguousarray(np.asarray(sbt, dtype=np.int64)).copy()
    invalid = L.orc_step_batch(kind, C.byref(params) if params is not None else None, n,
                               _ptr(st), _ptr(sbt_arr), _ptr(act), _ptr(obs), _ptr(reward),
                               _ptr(done))
    return dict(state=st, obs=obs, reward=reward, done=done, sbt=sbt_arr, invalid=int(invalid))


def reset_batch(kind, n, seed, global_env_offset=0, epoch=0, low=None, high=None, mask=None,
                state=None):
    L = lib()
    if state is None:
        st = np.zeros((STATE_DIM[kind], n), dtype=np.float64)
    else:
        st = np.ascontiguousarray(np.asarray(state, dtype=np.float64)).copy()
    lo = None if low is None else np.ascontiguousarray(np.asarray(low, dtype=np.float64))
    hi = None if high is None else np.ascontiguousarray(np.asarray(high, dtype=np.float64))
    m = None if mask is None else np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))
    L.orc_reset_batch(kind, n, _ptr(st), seed, global_env_offset, epoch, _ptr(lo), _ptr(hi),
                      _ptr(m))
    return st


def philox4x32_10(ctr, key):
    L = lib()
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    L.orc_philox4x32_10(c, k)
    return [int(x) for x in c]


def bench_regions(kind, n_envs, n_steps, n_warmup, n_burnin, n_regions, n_threads, seed=0):
    """The scalar reference loop over one set of env objects: n_burnin untimed steps, then n_regions
    regions of n_warmup untimed + n_steps timed steps.  Returns the list of region seconds."""
    L = lib()
    cs = C.c_double(0.0)
    out = (C.c_double * n_regions)()
    t = L.orc_bench_regions(kind, n_envs, n_steps, n_warmup, n_burnin, n_regions, n_threads, seed, out, C.byref(cs))
    if t < 0:
        raise MemoryError("orc_bench_regions: allocation failed")
    return [float(x) for x in out]


def bench_rollout(kind, n_envs, n_steps, n_warmup, n_threads, seed=0):
    """Time the scalar reference loop on host cores.  Returns (seconds, checksum)."""
    L = lib()
    cs = C.c_double(0.0)
    t = L.orc_bench_rollout(kind, n_envs, n_steps, n_warmup, n_threads, seed, C.byref(cs))
    return float(t), float(cs.value)
