"""LPIPS on the CUDA convolution kernels behind the reference's queued-metric contract.

Reference: utils/eval_metrics.py:100-156 -- ``PyIqaMetricFactory.get_metric('lpips')`` builds a ``BaseMetric`` whose
``calculate`` converts the grey frame to 3 channels (``cv2torch(num_ch=3)``, utils/eval_utils.py:46-54), queues it, and
every ``batch_size`` = 4 frames runs ``iqa_metric(img, ref)``; ``finish_queue`` flushes the remainder.  pyiqa and the
LPIPS weights it downloads are not part of the reference tree and do not exist offline, so the weights are supplied by
the caller (``state_dict`` with the lpips / pyiqa names, or a ``.pth`` path / ``EVREAL_LPIPS_WEIGHTS``); without them the
metric raises instead of inventing numbers.  Parity against pyiqa proper is UNPINNED; the CUDA
path is checked against oracle/metrics.py::lpips_oracle with seeded weights (tests/test_gpu_lpips.py).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .eval_metrics import BaseMetric, _cuda_img

BACKBONES = {'lpips': 0, 'lpips-alex': 0, 'alex': 0, 'lpips-vgg': 1, 'vgg': 1}
# torchvision `features` indices of the convolutions and the lpips slice they live in
_CONV_IDX = {0: (0, 3, 6, 8, 10), 1: (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)}


def _slice_of(backbone, idx):
    if backbone == 0:
        return 1 if idx < 2 else 2 if idx < 5 else 3 if idx < 8 else 4 if idx < 10 else 5
    return 1 if idx < 4 else 2 if idx < 9 else 3 if idx < 16 else 4 if idx < 23 else 5


def state_dict_from_conv_list(w, backbone):
    """{'conv{i}.weight|bias', 'lin{j}.weight'} (oracle/metrics.py naming) -> lpips state_dict names."""
    out = {}
    for i, idx in enumerate(_CONV_IDX[backbone]):
        pre = 'net.slice%d.%d' % (_slice_of(backbone, idx), idx)
        out[pre + '.weight'] = w['conv%d.weight' % i]
        out[pre + '.bias'] = w['conv%d.bias' % i]
    for j in range(5):
        out['lin%d.model.1.weight' % j] = w['lin%d.weight' % j]
    return out


_default_weights = None


def set_default_weights(state_dict):
    """Weights used by LpipsMetric instances created without any (the tracker builds its metrics by name only)."""
    global _default_weights
    _default_weights = state_dict


class LpipsNet:
    """evk_lpips handle: ``forward(img, ref)`` -> float64 scores [n] for n <= batch pairs of [H, W] frames in [0, 1]."""

    def __init__(self, backbone, state_dict, height, width, batch=4, precision=0):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.backbone = BACKBONES[backbone] if isinstance(backbone, str) else int(backbone)
        self.batch, self.H, self.W = int(batch), int(height), int(width)
        h = ctypes.c_void_p()
        _lib.check(self.lib.evk_lpips_create(self.backbone, self.batch, self.H, self.W, int(precision), ctypes.byref(h)))
        self.handle = h
        for name, t in state_dict.items():
            a = np.ascontiguousarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float32)
            shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
            _lib.check(self.lib.evk_lpips_load_tensor(self.handle, name.encode(), a.ctypes.data_as(ctypes.c_void_p), shape, a.ndim))
        _lib.check(self.lib.evk_lpips_finalize(self.handle))
        # kernels per forward: input scaling, 5 (AlexNet) / 13 (VGG16) convolutions, 3 / 4 max-poolings, 5 tap reductions, sum
        self.launches_per_forward = 1 + (5 + 3 if self.backbone == 0 else 13 + 4) + 5 + 1

    @property
    def num_tensor_core_layers(self):
        return int(self.lib.evk_lpips_num_tc_layers(self.handle))

    def forward(self, img, ref):
        a, b = _cuda_img(img), _cuda_img(ref)
        if a.dim() == 2:
            a, b = a[None], b[None]
        if a.shape != b.shape or tuple(a.shape[1:]) != (self.H, self.W) or not 1 <= a.shape[0] <= self.batch:
            raise ValueError("LpipsNet.forward: expected 1..%d pairs of %dx%d frames, got %s / %s"
                             % (self.batch, self.H, self.W, tuple(a.shape), tuple(b.shape)))
        out = torch.empty((a.shape[0],), dtype=torch.float64, device=a.device)
        with torch.cuda.device(a.device):
            _lib.check(self.lib.evk_lpips_forward(self.handle, _lib.ptr(a), _lib.ptr(b), int(a.shape[0]), _lib.ptr(out),
                                                  _lib.stream_ptr(a.device)))
        return out

    __call__ = forward

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.evk_lpips_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class LpipsMetric(BaseMetric):
    """Drop-in for the class PyIqaMetricFactory builds (utils/eval_metrics.py:118-156): queue of ``batch_size`` frames,
    ``calculate`` returns [] until the queue is full, ``finish_queue`` flushes."""

    def __init__(self, name='lpips', state_dict=None, weights_path=None, precision=0):
        super().__init__(name=name.lower(), no_ref=False)
        if name.lower() not in BACKBONES:
            raise ValueError("Unknown LPIPS variant: %s" % name)
        self.backbone = BACKBONES[name.lower()]
        self.precision = precision
        self._sd = state_dict
        self._path = weights_path or os.environ.get('EVREAL_LPIPS_WEIGHTS')
        self._net = None

    def _weights(self):
        if self._sd is not None:
            return self._sd
        if _default_weights is not None:
            return _default_weights
        if self._path and os.path.exists(self._path):
            sd = torch.load(self._path, map_location='cpu', weights_only=False)
            return sd.get('params', sd.get('state_dict', sd)) if isinstance(sd, dict) else sd
        raise _lib.EvkError(
            "LPIPS weights are not available: the reference downloads them through pyiqa; pass state_dict=..., "
            "weights_path=... or set EVREAL_LPIPS_WEIGHTS to a .pth with the lpips/pyiqa state_dict")

    def _ensure(self, H, W):
        if self._net is None or (self._net.H, self._net.W) != (H, W):
            self._net = LpipsNet(self.backbone, self._weights(), H, W, self.batch_size, self.precision)
        return self._net

    def inference(self):
        if len(self.image_queue) < 1:
            return []
        imgs = torch.stack(self.image_queue[-self.batch_size:])
        refs = torch.stack(self.ref_queue[-self.batch_size:])
        net = self._ensure(int(imgs.shape[1]), int(imgs.shape[2]))
        scores = net(imgs, refs).tolist()
        self.image_queue = []
        self.ref_queue = []
        return scores

    def finish_queue(self):
        self.updated = 0
        score = self.inference()
        self.updated += len(score)
        self.scores.extend(score)

    def calculate(self, img, ref=None):
        if ref is None:
            raise ValueError("LPIPS is a full-reference metric")
        self.image_queue.append(_cuda_img(img))
        self.ref_queue.append(_cuda_img(ref))
        if len(self.image_queue) < self.batch_size:
            return []
        return self.inference()
