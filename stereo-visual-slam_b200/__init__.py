"""stereo-visual-slam_b200: B200-native stereo-VO + sliding-window-BA hot path.

Host-side Python mirror of the reference's stage functions over the C-ABI in include/vslam_b200.h.
The directory name carries a hyphen (it mirrors the reference repo name); import it through the
root-level `vslam_b200_loader` shim, which registers it as `stereo_visual_slam_b200`.
"""
from . import ffi, sharding, synth  # noqa: F401
from .ffi import Context, VslamError, KEYPOINT_DTYPE, DMATCH_DTYPE  # noqa: F401
