"""gym_rs::envs::classical_control::cartpole (reference: src/envs/classical_control/cartpole.rs)."""
from __future__ import annotations

from dataclasses import dataclass

from ... import _capi
from ...core import Env, Metadata
from ...utils.renderer import RenderMode


@dataclass(frozen=True)
class CartPoleObservation:
    """cartpole.rs:327-334; Vec<f64>::from order is [x, x_dot, theta, theta_dot] (:336-349)."""
    x: float
    x_dot: float
    theta: float
    theta_dot: float

    def __neg__(self):  # cartpole.rs:367-378
        return CartPoleObservation(-self.x, -self.x_dot, -self.theta, -self.theta_dot)

    def to_vec(self):
        return [self.x, self.x_dot, self.theta, self.theta_dot]


class KinematicsIntegrator:
    """cartpole.rs:380-387"""
    Euler = 0
    Other = 1


class CartPoleEnv(Env):
    """CartPoleEnv::new(render_mode) (cartpole.rs:91-144) with `num_envs` instances on one GPU.

    Dynamics follow the reference, including `polemass_length = masspole + length`
    (cartpole.rs:150-152) and the absence of truncation (:480)."""
    KIND = _capi.CARTPOLE
    OBSERVATION = CartPoleObservation
    STATE = CartPoleObservation
    WARNS_AFTER_TERMINATION = True
    ACTION_DTYPE = "int32"
    INFO_ON_STEP = ()                    # info: Some(()), cartpole.rs:481
    INVALID_FMT = "{} usize invalid"     # cartpole.rs:404
    _METADATA = Metadata((RenderMode.Human, RenderMode.RgbArray), 50)  # cartpole.rs:264-270

    @property
    def steps_beyond_terminated(self):
        """cartpole.rs:81: None, or Some(k).  num_envs == 1 -> Optional[int]; else an int32 device
        view with -1 for None."""
        if self.num_envs == 1:
            _, sbt = self.get_state(with_sbt=True)
            return None if sbt[0] < 0 else int(sbt[0])
        self.sync()
        return self._t_sbt
