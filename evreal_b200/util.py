"""Host-side mirror of EVREAL's ``utils/util.py`` and of ``eval.normalize_event_tensor``.

``CropParameters`` keeps the reference's attributes (utils/util.py:30-59); pad and
crop run as CUDA kernels when given CUDA tensors.
"""
import json
from collections import OrderedDict
from math import ceil, floor
from pathlib import Path

import torch

from . import _lib


def read_json(fname):
    with Path(fname).open('rt', encoding="utf-8") as handle:
        return json.load(handle, object_hook=OrderedDict)


def get_height_width(data_loader):
    for d in data_loader:
        return d['events'].shape[-2:]


def optimal_crop_size(max_size, max_subsample_factor, safety_margin=0):
    """Smallest size >= max_size divisible by 2**max_subsample_factor (utils/util.py:19-27)."""
    m = pow(2, max_subsample_factor)
    return int(m * ceil(max_size / m)) + safety_margin * m


def _as4d(x):
    shape = x.shape
    if x.dim() < 2:
        raise ValueError("expected a tensor with at least 2 dimensions")
    H, W = shape[-2], shape[-1]
    planes = 1
    for s in shape[:-2]:
        planes *= s
    return x.reshape(1, planes, H, W), shape


def normalize_pad(voxel, Hp, Wp, normalize):
    """normalize_event_tensor (eval.py:398-410, per sample) fused with CropParameters.pad.
    voxel: CUDA float32 [N, C, H, W] -> [N, C, Hp, Wp]."""
    lib = _lib.load()
    assert voxel.is_cuda and voxel.dtype == torch.float32 and voxel.dim() == 4
    voxel = voxel.contiguous()
    N, C, H, W = voxel.shape
    out = torch.empty((N, C, Hp, Wp), dtype=torch.float32, device=voxel.device)
    with torch.cuda.device(voxel.device):
        _lib.check(lib.evk_normalize_pad(_lib.ptr(voxel), _lib.ptr(out), N, C, H, W, Hp, Wp, int(bool(normalize)),
                                         _lib.stream_ptr(voxel.device)))
    return out


def normalize_event_tensor(event_tensor):
    """Mean/std normalisation over the non-zero entries (eval.py:398-410).

    The reference normalises the whole tensor at once and is only ever called
    with batch 1; with a batch this normalises per sample (the only
    parity-safe way to batch)."""
    _lib.require_cuda()
    x = event_tensor if event_tensor.is_cuda else event_tensor.cuda(non_blocking=True)
    x = x.float()
    shape = x.shape
    if x.dim() == 4:
        x4 = x
    else:
        x4 = x.reshape(1, -1, shape[-2], shape[-1])
    return normalize_pad(x4, shape[-2], shape[-1], True).reshape(shape)


class CropParameters:
    """Pad to a multiple of 2**num_encoders and centre-crop back (utils/util.py:30-59)."""

    def __init__(self, width, height, num_encoders, safety_margin=0):
        self.height = height
        self.width = width
        self.num_encoders = num_encoders
        self.width_crop_size = optimal_crop_size(self.width, num_encoders, safety_margin)
        self.height_crop_size = optimal_crop_size(self.height, num_encoders, safety_margin)
        self.padding_top = ceil(0.5 * (self.height_crop_size - self.height))
        self.padding_bottom = floor(0.5 * (self.height_crop_size - self.height))
        self.padding_left = ceil(0.5 * (self.width_crop_size - self.width))
        self.padding_right = floor(0.5 * (self.width_crop_size - self.width))
        self.cx = floor(self.width_crop_size / 2)
        self.cy = floor(self.height_crop_size / 2)
        self.ix0 = self.cx - floor(self.width / 2)
        self.ix1 = self.cx + ceil(self.width / 2)
        self.iy0 = self.cy - floor(self.height / 2)
        self.iy1 = self.cy + ceil(self.height / 2)

    def pad(self, x):
        if not x.is_cuda:
            _lib.require_cuda()
            x = x.cuda(non_blocking=True)
        x4, shape = _as4d(x.float())
        out = normalize_pad(x4, self.height_crop_size, self.width_crop_size, False)
        return out.reshape(tuple(shape[:-2]) + (self.height_crop_size, self.width_crop_size))

    def crop(self, img):
        if not img.is_cuda:
            _lib.require_cuda()
            img = img.cuda(non_blocking=True)
        lib = _lib.load()
        x4, shape = _as4d(img.float().contiguous())
        H, W = self.iy1 - self.iy0, self.ix1 - self.ix0
        out = torch.empty((1, x4.shape[1], H, W), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            _lib.check(lib.evk_crop(_lib.ptr(x4), _lib.ptr(out), 1, x4.shape[1], shape[-2], shape[-1], H, W,
                                    _lib.stream_ptr(img.device)))
        return out.reshape(tuple(shape[:-2]) + (H, W))
