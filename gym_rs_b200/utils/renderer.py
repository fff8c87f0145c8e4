"""gym_rs::utils::renderer -- only the types the step path names.  SDL2 rendering is out of scope
(BASELINE.json north_star); every env here behaves like RenderMode::None, under which the
reference's renderer is a no-op (src/utils/renderer.rs:40-62)."""
import enum


class RenderMode(enum.Enum):
    """src/utils/renderer.rs:83-114"""
    Human = "human"
    SingleRgbArray = "single_rgb_array"
    RgbArray = "rgb_array"
    DepthArray = "depth_array"
    SingleDepthArray = "single_depth_array"
    NONE = "none"


class Renders(enum.Enum):
    """src/utils/renderer.rs:118-130"""
    NONE = "none"
