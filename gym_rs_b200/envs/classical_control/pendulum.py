"""Pendulum-v1.  NOT in the reference crate (src/envs/classical_control/mod.rs:1-4 lists only
cartpole and mountain_car); BASELINE.json asks for it, so it follows upstream OpenAI Gym
pendulum.py with gym-rs conventions (f32 device / f64 oracle, done = truncated = false, info None).
Parity for this env is unpinned by the reference."""
from __future__ import annotations

from dataclasses import dataclass

from ... import _capi
from ...core import Env, Metadata
from ...utils.renderer import RenderMode


@dataclass(frozen=True)
class PendulumObservation:
    cos_theta: float
    sin_theta: float
    theta_dot: float

    def to_vec(self):
        return [self.cos_theta, self.sin_theta, self.theta_dot]


@dataclass(frozen=True)
class PendulumState:
    theta: float
    theta_dot: float


class PendulumEnv(Env):
    KIND = _capi.PENDULUM
    OBSERVATION = PendulumObservation
    STATE = PendulumState
    ACTION_DTYPE = "float32"
    INFO_ON_STEP = None
    INVALID_FMT = "{} invalid"
    _METADATA = Metadata((RenderMode.Human, RenderMode.RgbArray), 30)
