"""Oracle: event windows -> voxel grids (stage 1) and the glue around it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, in the reference tree (/root/reference):
  utils/event_utils.py:4-24    events_to_image_torch  (scatter-add, accumulate=True)
  utils/event_utils.py:27-59   events_to_voxel_torch  (bilinear-in-time weights per bin)
  dataset.py:222-228, 52-58    get_events / __getitem__ casts (raw arrays -> f32 tensors)
  eval.py:398-410              normalize_event_tensor
  utils/util.py:19-59          optimal_crop_size / CropParameters
"""
from math import ceil, floor

import numpy as np
import torch


def linspace_f32(start, end, steps):
    """torch.linspace(start, end, steps) for float32 on CPU, restated.

    ATen computes ``step = (end-start)/(steps-1)`` in float32 and fills the first
    half forward from ``start`` and the second half backward from ``end``.
    Checked element-for-element against torch in tests/test_oracle_voxel.py.
    """
    start = np.float32(start)
    end = np.float32(end)
    if steps == 1:
        return np.array([start], dtype=np.float32)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps, dtype=np.int64)
    half = steps // 2
    fwd = (start + step * i.astype(np.float32)).astype(np.float32)
    bwd = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, fwd, bwd).astype(np.float32)


def t_norm_f32(ts, num_bins):
    """Normalised event time in [0, num_bins-1] -- utils/event_utils.py:47-51."""
    ts = np.asarray(ts, dtype=np.float32)
    dt = np.float32(ts[-1] - ts[0])
    if float(dt) < 1e-9:
        return linspace_f32(0.0, num_bins - 1, len(ts))
    return ((ts - ts[0]) / dt * np.float32(num_bins - 1)).astype(np.float32)


def events_to_voxel_numpy(xs, ys, ts, ps, num_bins, sensor_size=(180, 240)):
    """Sequential float32 restatement (events added in index order, one bin at a time).

    Matches the reference bit-for-bit whenever ATen's index_put_ runs single
    threaded (N <~ 60k); for larger N the reference itself is only
    reproducible to ~3e-6.
    Raises IndexError on out-of-range coordinates like the reference does;
    negative coordinates wrap (python indexing) like the reference.
    """
    xs = np.asarray(xs, dtype=np.float32)
    ys = np.asarray(ys, dtype=np.float32)
    ps = np.asarray(ps, dtype=np.float32)
    assert len(xs) == len(ys) == len(ts) == len(ps)
    H, W = sensor_size
    xi = xs.astype(np.int64)
    yi = ys.astype(np.int64)
    if len(xi) and (xi.max() >= W or yi.max() >= H or xi.min() < -W or yi.min() < -H):
        raise IndexError("event coordinate out of range for sensor_size %s" % (sensor_size,))
    tn = t_norm_f32(ts, num_bins)
    out = np.zeros((num_bins, H, W), dtype=np.float32)
    one = np.float32(1.0)
    for b in range(num_bins):
        w = np.maximum(np.float32(0.0), one - np.abs(tn - np.float32(b))).astype(np.float32)
        np.add.at(out[b], (yi, xi), (ps * w).astype(np.float32))
    return out


def events_to_voxel_oracle(xs, ys, ts, ps, num_bins, sensor_size=(180, 240)):
    """Same arithmetic with torch CPU ops (the cost model of the reference: one
    index_put_(accumulate=True) pass over all events per bin).  Used as the
    timed CPU baseline; returns a torch tensor [num_bins, H, W]."""
    assert len(xs) == len(ys) == len(ts) == len(ps)
    dt = ts[-1] - ts[0]
    if dt.item() < 1e-9:
        tn = torch.linspace(0, num_bins - 1, steps=len(ts))
    else:
        tn = (ts - ts[0]) / dt * (num_bins - 1)
    yi = ys.long()
    xi = xs.long()
    planes = torch.zeros((num_bins,) + tuple(sensor_size), dtype=torch.float32)
    for b in range(num_bins):
        w = torch.clamp_min(1.0 - (tn - b).abs(), 0.0)
        planes[b].index_put_((yi, xi), ps * w, accumulate=True)
    return planes


def raw_window_to_f32(xy, t, p):
    """Raw on-disk slices -> the four f32 arrays the voxelizer takes.
    dataset.py:222-228 (xs, ys f32; ps = p*2.0-1.0) and :52-58 (ts - ts[0] in
    f64, then rounded to f32)."""
    xy = np.asarray(xy)
    xs = xy[:, 0].astype(np.float32)
    ys = xy[:, 1].astype(np.float32)
    t = np.asarray(t, dtype=np.float64)
    ps = (np.asarray(p) * 2.0 - 1.0).astype(np.float32)
    ts = (t - t[0]).astype(np.float32) if len(t) else t.astype(np.float32)
    return xs, ys, ts, ps


def normalize_event_tensor_oracle(v):
    """eval.py:398-410.  ``v`` is a torch tensor; statistics over nonzero entries."""
    nz = v != 0
    n = nz.sum()
    if n > 0:
        mean = v.sum() / n
        std = torch.sqrt((v ** 2).sum() / n - mean ** 2)
        std = torch.max(std, torch.tensor(1e-6))
        v = nz.float() * (v - mean) / std
    return v


class CropOracle:
    """utils/util.py:30-59 -- zero-pad to a multiple of 2**num_encoders, centre-crop back."""

    def __init__(self, width, height, num_encoders):
        m = 2 ** num_encoders
        self.width, self.height = width, height
        self.wp = int(m * ceil(width / m))
        self.hp = int(m * ceil(height / m))
        self.top = ceil(0.5 * (self.hp - height))
        self.bottom = floor(0.5 * (self.hp - height))
        self.left = ceil(0.5 * (self.wp - width))
        self.right = floor(0.5 * (self.wp - width))
        cx, cy = floor(self.wp / 2), floor(self.hp / 2)
        self.ix0, self.ix1 = cx - floor(width / 2), cx + ceil(width / 2)
        self.iy0, self.iy1 = cy - floor(height / 2), cy + ceil(height / 2)

    def pad(self, x):
        return torch.nn.functional.pad(x, (self.left, self.right, self.top, self.bottom))

    def crop(self, x):
        return x[..., self.iy0:self.iy1, self.ix0:self.ix1]
