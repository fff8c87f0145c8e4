"""gym_rs::utils::custom::util_fns (reference: src/utils/custom/util_fns.rs:2-10)."""
from ... import _capi


def clip(value, left_bound, right_bound):
    """Same branch order as the reference: in range -> value; > right -> right; else left."""
    if all(isinstance(v, int) for v in (value, left_bound, right_bound)):
        # the generic function on integers, as the reference's own tests call it
        return int(_capi.load().gymrs_clip(float(value), float(left_bound), float(right_bound)))
    return _capi.load().gymrs_clip(value, left_bound, right_bound)
