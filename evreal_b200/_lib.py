"""ctypes binding of libevreal_b200.so (the C ABI declared in include/evreal_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA
device is present, importing/using the ops raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libevreal_b200.so")

EVK_OK, EVK_ERR_ARG, EVK_ERR_CUDA, EVK_ERR_INDEX, EVK_ERR_STATE, EVK_ERR_KEY = 0, -1, -2, -3, -4, -5


class EventWindow(ctypes.Structure):
    """evk_event_window: one raw event window (device or pinned-host pointers)."""
    _fields_ = [("xy", ctypes.c_void_p), ("t", ctypes.c_void_p), ("pol", ctypes.c_void_p), ("n", ctypes.c_int64)]


class ModelConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "arch", "num_bins", "base_channels", "num_encoders", "num_residual_blocks", "kernel_size",
        "num_output_channels", "final_sigmoid", "dynamic_decoder", "batch", "height", "width", "precision")]


_c = ctypes
_vp, _i, _i64, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double
# name -> (restype, argtypes); kept in sync with include/evreal_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "evk_version": (_i, []),
    "evk_last_error": (_c.c_char_p, []),
    "evk_voxelize": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "evk_voxelize_raw": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "evk_voxelize_raw_batch": (_i, [_c.POINTER(EventWindow), _i, _i, _i, _i, _vp, _vp, _vp]),
    "evk_stage_windows_h2d": (_i, [_c.POINTER(EventWindow), _i, _vp, _vp, _vp, _i64, _vp]),
    "evk_stage_frames_h2d": (_i, [_c.POINTER(_vp), _i, _i64, _vp, _vp]),
    "evk_u8_to_f32_batch": (_i, [_c.POINTER(_vp), _i, _i64, _vp, _vp]),
    "evk_normalize_pad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "evk_model_create": (_i, [_c.POINTER(ModelConfig), _c.POINTER(_vp)]),
    "evk_model_load_tensor": (_i, [_vp, _c.c_char_p, _vp, _c.POINTER(_i64), _i]),
    "evk_model_finalize": (_i, [_vp, _vp]),
    "evk_model_reset_states": (_i, [_vp, _vp]),
    "evk_model_forward": (_i, [_vp, _vp, _vp, _vp]),
    "evk_model_num_states": (_i, [_vp]),
    "evk_model_state_shape": (_i, [_vp, _i, _c.POINTER(_i64)]),
    "evk_model_get_state": (_i, [_vp, _i, _vp, _vp]),
    "evk_model_set_state": (_i, [_vp, _i, _vp, _vp]),
    "evk_model_destroy": (_i, [_vp]),
    "evk_model_io_buffers": (_i, [_vp, _c.POINTER(_vp), _c.POINTER(_vp)]),
    "evk_model_last_launch_count": (_i, [_vp]),
    "evk_model_flops": (_d, [_vp]),
    "evk_model_profile": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _c.POINTER(_i)]),
    "evk_model_op_desc": (_i, [_vp, _i, _c.c_char_p, _i]),
    "evk_conv2d_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "evk_pack_layer_weights": (_i, [_i, _vp, _i, _i, _i, _i, _i, _vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "evk_percentile_normalize": (_i, [_vp, _vp, _i, _i, _d, _d, _i, _vp]),
    "evk_crop": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "evk_mse_ssim": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "evk_u8_to_f32": (_i, [_vp, _vp, _i64, _vp]),
    "evk_quantize_u8": (_i, [_vp, _vp, _i64, _vp]),
    "evk_searchsorted_f64": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _vp]),
    "evk_equalize_hist": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "evk_equalize_local": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "evk_lpips_create": (_i, [_i, _i, _i, _i, _i, _c.POINTER(_vp)]),
    "evk_lpips_load_tensor": (_i, [_vp, _c.c_char_p, _vp, _c.POINTER(_i64), _i]),
    "evk_lpips_finalize": (_i, [_vp]),
    "evk_lpips_forward": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "evk_lpips_flops": (_d, [_vp]),
    "evk_lpips_num_tc_layers": (_i, [_vp]),
    "evk_lpips_destroy": (_i, [_vp]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "evreal_b200: %s is missing -- build it with `python -m evreal_b200.build` "
            "(there is no CPU fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EvkError(RuntimeError):
    pass


def check(rc):
    """Map an EVK_ERR_* return code to the exception type the reference would raise."""
    if rc == EVK_OK:
        return
    msg = load().evk_last_error().decode("utf-8", "replace")
    if rc == EVK_ERR_ARG:
        raise ValueError(msg)
    if rc == EVK_ERR_INDEX:
        raise IndexError(msg)
    if rc == EVK_ERR_KEY:
        raise KeyError(msg)
    raise EvkError("evreal_b200 error %d: %s" % (rc, msg))


def require_cuda():
    if not torch.cuda.is_available():
        raise EvkError("evreal_b200 needs a CUDA device (sm_100a); there is no CPU implementation of the hot path")


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())
