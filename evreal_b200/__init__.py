"""evreal_b200 -- B200-native hot path of EVREAL's eval loop.

voxelizer -> recurrent reconstruction network -> MSE / SSIM (/ LPIPS), behind the
reference's own plugin surface.  The arithmetic lives in ``libevreal_b200.so``
(hand-written sm_100a CUDA behind the C ABI of ``include/evreal_b200.h``); the
modules here mirror the reference's Python interfaces.  There is no CPU
fallback: using any op without the built library or without a CUDA device raises.
"""
from . import _lib  # noqa: F401
from .event_utils import events_to_voxel_torch, events_to_image_torch, events_to_voxel_raw  # noqa: F401
from .model import E2VIDRecurrent, FlowNet, FireNet, FireNet_legacy, SpadeE2vid, EITR, ColorNet  # noqa: F401
from .util import CropParameters, normalize_event_tensor  # noqa: F401

__version__ = "0.1.0"
