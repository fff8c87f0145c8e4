"""gym_rs::utils::custom"""
from . import util_fns  # noqa: F401
