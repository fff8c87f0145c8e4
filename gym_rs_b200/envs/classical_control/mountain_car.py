"""gym_rs::envs::classical_control::mountain_car
(reference: src/envs/classical_control/mountain_car.rs)."""
from __future__ import annotations

from dataclasses import dataclass

from ... import _capi
from ...core import Env, Metadata
from ...utils.renderer import RenderMode


@dataclass(frozen=True, order=True)
class MountainCarObservation:
    """mountain_car.rs:121-128; Vec<f64>::from order is [position, velocity] (:193-197)."""
    position: float
    velocity: float

    def to_vec(self):
        return [self.position, self.velocity]


class MountainCarEnv(Env):
    """MountainCarEnv::new(render_mode) (mountain_car.rs:341-389) with `num_envs` instances."""
    KIND = _capi.MOUNTAIN_CAR
    OBSERVATION = MountainCarObservation
    STATE = MountainCarObservation
    ACTION_DTYPE = "int32"
    INFO_ON_STEP = None                  # info: None, mountain_car.rs:433
    INVALID_FMT = "{} (usize) invalid"   # mountain_car.rs:404
    _METADATA = Metadata((RenderMode.Human, RenderMode.RgbArray, RenderMode.SingleRgbArray,
                          RenderMode.NONE), 30)  # mountain_car.rs:108-119
