"""gym_rs::envs::classical_control (reference: src/envs/classical_control/mod.rs) + Pendulum."""
from . import cartpole, mountain_car, pendulum  # noqa: F401
