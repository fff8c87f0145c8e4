"""gym_rs_b200 -- B200-native batched classic-control env stepper behind gym-rs's Env surface.

Module paths follow the reference crate (`gym_rs::core`, `gym_rs::spaces`,
`gym_rs::envs::classical_control::{cartpole, mountain_car}`, `gym_rs::utils::...`).
All compute happens in hand-written sm_100a kernels reached through the C ABI in
include/gymrs_b200.h (gym_rs_b200/csrc); this Python layer is a thin ctypes host binding
used by the tests and bench.  There is no CPU implementation in this package.
"""
from . import core, spaces  # noqa: F401
from .core import ActionReward, Env, EnvProperties, RewardRange  # noqa: F401
from .envs.classical_control.cartpole import CartPoleEnv, CartPoleObservation  # noqa: F401
from .envs.classical_control.mountain_car import MountainCarEnv, MountainCarObservation  # noqa: F401
from .envs.classical_control.pendulum import PendulumEnv, PendulumObservation  # noqa: F401
from .spaces import BoxR, Discrete  # noqa: F401
from .utils.renderer import RenderMode  # noqa: F401

__all__ = ["core", "spaces", "ActionReward", "Env", "EnvProperties", "RewardRange", "CartPoleEnv",
           "CartPoleObservation", "MountainCarEnv", "MountainCarObservation", "PendulumEnv",
           "PendulumObservation", "BoxR", "Discrete", "RenderMode"]
