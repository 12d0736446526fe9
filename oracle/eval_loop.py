"""Oracle: the per-sequence evaluation loop on the CPU (the reference's eval.py:189-246 path).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Also the timed CPU baseline /
``bench.py --impl reference`` arm: it executes the same torch-CPU operators the
reference executes per frame (index_put_ voxelizer, F.conv2d network, numpy
percentile + scipy-style SSIM), in the same order, without the reference's file
outputs (save_images=False there too).

Follows, in /root/reference:
  eval.py:189-246            eval_method_on_sequence (skip / break rules, normalisation, pad, crop, post-norm)
  dataset.py:33-102          MemMapDataset.__getitem__ ('between_frames': item i = events between frame i-1 and i)
  utils/eval_metrics.py:244-273  clip + gating (evaluation window, timestamp tolerance)
"""
import time

import numpy as np
import torch

from . import event_voxel as ov
from . import metrics as om
from .windows import between_frames_windows


def run_sequence(arrays, sensor_resolution, model, num_encoders, event_tensor_normalization, post_process_norm,
                 start_time_s=None, end_time_s=None, ts_tol_ms=1.0, num_bins=5, max_items=None, timers=None,
                 compute_metrics=True):
    """arrays: dict with events_ts / events_xy / events_p / images / images_ts / image_event_indices (on-disk layout).
    Returns dict(indices, mse, ssim, frames, events, images)."""
    H, W = int(sensor_resolution[0]), int(sensor_resolution[1])
    t_all, xy_all, p_all = arrays['events_ts'], arrays['events_xy'], arrays['events_p']
    frame_ts = [float(v) for v in np.asarray(arrays['images_ts']).reshape(-1)]
    windows = between_frames_windows(arrays['image_event_indices'])
    if start_time_s is None:
        start_time_s = min(frame_ts[0], float(t_all[0]))
    if end_time_s is None:
        end_time_s = max(frame_ts[-1], float(t_all[-1]))
    crop = ov.CropOracle(W, H, num_encoders)
    model.reset_states()
    out = {'indices': [], 'mse': [], 'ssim': [], 'frames': 0, 'events': 0, 'images': []}
    tm = timers if timers is not None else {}
    for k in ('voxel', 'model', 'metrics'):
        tm.setdefault(k, 0.0)
    for idx, (i0, i1) in enumerate(windows):
        if max_items is not None and idx >= max_items:
            break
        ts_frame = frame_ts[idx]                              # between_frames: voxel_timestamp == frame_timestamp
        if ts_frame < start_time_s - 10:
            continue
        if ts_frame > end_time_s:
            break
        t0 = time.perf_counter()
        if i1 > i0:
            xs, ys, ts, ps = ov.raw_window_to_f32(xy_all[i0:i1], t_all[i0:i1], p_all[i0:i1])
            voxel = ov.events_to_voxel_oracle(torch.from_numpy(xs), torch.from_numpy(ys), torch.from_numpy(ts),
                                              torch.from_numpy(ps), num_bins, (H, W))
        else:
            voxel = torch.zeros((num_bins, H, W), dtype=torch.float32)
        voxel = voxel[None]
        if event_tensor_normalization:
            voxel = ov.normalize_event_tensor_oracle(voxel)
        voxel = crop.pad(voxel)
        t1 = time.perf_counter()
        image = crop.crop(model(voxel))
        t2 = time.perf_counter()
        image = om.post_process_oracle(image[0, 0].numpy(), post_process_norm)
        image = np.clip(image, 0.0, 1.0)
        ref = np.clip(arrays['images'][idx][:, :, 0].astype(np.float32) / 255, 0.0, 1.0)
        out['images'].append(image)
        inside = start_time_s <= ts_frame <= end_time_s      # img_ts == ref_ts -> tolerance always met
        if inside and compute_metrics:
            out['indices'].append(idx)
            out['mse'].append(om.mse_oracle(image, ref))
            out['ssim'].append(om.ssim_oracle(image, ref))
        t3 = time.perf_counter()
        tm['voxel'] += t1 - t0
        tm['model'] += t2 - t1
        tm['metrics'] += t3 - t2
        out['frames'] += 1
        out['events'] += max(i1 - i0, 0)
    return out


def weighted_means(per_sequence):
    """eval.py:259-266,367-368: dataset mean = sum(mean_seq * n_seq) / sum(n_seq); a sequence without scores
    contributes its -1 placeholder with weight 0 (skipped because count == 0)."""
    tot = {}
    for n_eval, means in per_sequence:
        if n_eval == 0:
            continue
        for k, v in means.items():
            a = tot.setdefault(k, [0.0, 0])
            a[0] += v * n_eval
            a[1] += n_eval
    return {k: a[0] / a[1] for k, a in tot.items()}
