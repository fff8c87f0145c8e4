"""gym_rs::utils"""
from . import renderer, seeding  # noqa: F401
