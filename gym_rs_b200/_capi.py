"""ctypes declarations for include/gymrs_b200.h (the C ABI of libgymrs_b200.so).

Loading never falls back to anything: if the shared library is missing it is built with
nvcc (gym_rs_b200.build), and if that fails the import error propagates.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

OK, ERR_INVALID_ACTION, ERR_BAD_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC, ERR_UNSUPPORTED = range(7)
CARTPOLE, MOUNTAIN_CAR, PENDULUM = 0, 1, 2
FLAG_TIME_LIMIT = 0x1
STEP_AUTORESET = 0x1
HOST_U8_ACTIONS, HOST_PACKED_DONE = 0x1, 0x2


class CartPoleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "gravity", "masscart", "masspole", "length", "force_mag", "tau",
        "theta_threshold_radians", "x_threshold")] + [
        ("kinematics_integrator", C.c_int32), ("max_episode_steps", C.c_int32)]


class MountainCarParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "min_position", "max_position", "max_speed", "goal_position", "goal_velocity",
        "force", "gravity")] + [("max_episode_steps", C.c_int32), ("_pad", C.c_int32)]


class PendulumParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("max_speed", "max_torque", "dt", "g", "m", "l")] + [
        ("max_episode_steps", C.c_int32), ("_pad", C.c_int32)]


PARAMS = {CARTPOLE: CartPoleParams, MOUNTAIN_CAR: MountainCarParams, PENDULUM: PendulumParams}


class Buffers(C.Structure):
    _fields_ = [("num_envs", C.c_uint64), ("ld", C.c_uint64), ("state_dim", C.c_uint32),
                ("obs_dim", C.c_uint32), ("state", C.c_void_p), ("obs", C.c_void_p),
                ("reward", C.c_void_p), ("done", C.c_void_p), ("truncated", C.c_void_p),
                ("steps_beyond_terminated", C.c_void_p), ("elapsed_steps", C.c_void_p)]


class CheckpointInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_uint32), ("num_envs", C.c_uint64),
                ("global_env_offset", C.c_uint64), ("seed", C.c_uint64), ("step_count", C.c_uint64),
                ("bytes", C.c_uint64)]


HOST_STEP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32)


class HostRolloutDesc(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p),
                ("truncated", C.c_void_p), ("action_slots", C.c_uint32), ("result_slots", C.c_uint32),
                ("transport", C.c_uint32), ("_pad", C.c_uint32), ("on_step", HOST_STEP_FN), ("user", C.c_void_p)]


_vp, _u64, _u32, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
_pu64 = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); exactly the functions include/gymrs_b200.h declares
SIGNATURES = {
    "gymrs_abi_version": (_i, []),
    "gymrs_last_error": (C.c_char_p, []),
    "gymrs_device_count": (_i, []),
    "gymrs_default_params": (_i, [_i, _vp]),
    "gymrs_create": (_i, [_i, _u64, _i, _u64, _vp, _u32, C.POINTER(_vp)]),
    "gymrs_destroy": (_i, [_vp]),
    "gymrs_clone": (_i, [_vp, C.POINTER(_vp)]),
    "gymrs_set_params": (_i, [_vp, _vp]),
    "gymrs_get_params": (_i, [_vp, _vp]),
    "gymrs_set_stream": (_i, [_vp, _vp]),
    "gymrs_get_stream": (_i, [_vp, C.POINTER(_vp)]),
    "gymrs_reset": (_i, [_vp, _pu64, _vp, _vp, _vp, _pu64]),
    "gymrs_step": (_i, [_vp, _vp, _u32]),
    "gymrs_step_many": (_i, [C.POINTER(_vp), C.POINTER(_vp), _u32, _u32, C.POINTER(_u32)]),
    "gymrs_step_host": (_i, [_vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "gymrs_step_host_async": (_i, [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _pu64]),
    "gymrs_host_wait": (_i, [_vp, _u64]),
    "gymrs_rollout_host": (_i, [_vp, _u32, _u32, C.POINTER(HostRolloutDesc)]),
    "gymrs_rollout": (_i, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "gymrs_get_state": (_i, [_vp, _vp, _vp]),
    "gymrs_set_state": (_i, [_vp, _vp, _vp]),
    "gymrs_get_buffers": (_i, [_vp, C.POINTER(Buffers)]),
    "gymrs_checkpoint_size": (_i, [_vp, C.POINTER(C.c_size_t)]),
    "gymrs_checkpoint_save": (_i, [_vp, _vp, C.c_size_t]),
    "gymrs_checkpoint_load": (_i, [_vp, _vp, C.c_size_t]),
    "gymrs_checkpoint_create": (_i, [_vp, C.c_size_t, _i, C.POINTER(_vp)]),
    "gymrs_checkpoint_info_of": (_i, [_vp, C.c_size_t, C.POINTER(CheckpointInfo)]),
    "gymrs_action_space": (_i, [_vp, _pu64, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "gymrs_observation_space": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gymrs_reward_range": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gymrs_num_envs": (_i, [_vp, _pu64]),
    "gymrs_kind_of": (_i, [_vp, C.POINTER(_i)]),
    "gymrs_sync": (_i, [_vp, _pu64]),
    "gymrs_set_launch_config": (_i, [_vp, _i, _i, _i]),
    "gymrs_set_launch_occupancy": (_i, [_vp, _i]),
    "gymrs_step_pass": (_i, [C.POINTER(_vp), C.POINTER(_vp), _u32, _u32, _vp, _vp, C.POINTER(_u32)]),
    "gymrs_host_alloc": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "gymrs_host_free": (_i, [_vp]),
    "gymrs_clip": (C.c_double, [C.c_double, C.c_double, C.c_double]),
    "gymrs_discrete_contains": (_i, [_u64, _u64]),
    "gymrs_rand_random": (_u64, [_pu64]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if needed) libgymrs_b200.so.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    # GYMRS_LIB_PATH: an experimental build of the same sources (tools/build_variant.py), for A/B runs
    path = os.environ.get("GYMRS_LIB_PATH") or _build.build()
    if not os.path.exists(path):
        raise ImportError("libgymrs_b200.so is missing and could not be built")
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if L.gymrs_abi_version() != 1:
        raise ImportError("libgymrs_b200.so ABI version mismatch")
    _lib = L
    return L


class GymrsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gymrs error {code}: {msg}")
        self.code = code


def check(code: int) -> None:
    if code != OK:
        raise GymrsError(code, load().gymrs_last_error().decode(errors="replace"))
