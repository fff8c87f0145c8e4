"""gym_rs::utils::seeding (reference: src/utils/seeding.rs:21-26)."""
import ctypes as C

from .. import _capi


def rand_random(seed=None):
    """Returns (generator_key, seed_no).  The reference returns a PCG64 seeded with seed_no; here
    the generator is the counter-based Philox4x32-10 keyed by seed_no, so the key IS the state."""
    L = _capi.load()
    if seed is None:
        seed_no = L.gymrs_rand_random(None)
    else:
        s = C.c_uint64(seed)
        seed_no = L.gymrs_rand_random(C.byref(s))
    return ("philox4x32_10", seed_no), seed_no
