"""Deterministic synthetic event streams in EVREAL's on-disk sequence format.

Format (reference tools/bag_to_npy.py:54-94, dataset.py:230-250):
``events_ts.npy`` f64 [N] sorted seconds, ``events_xy.npy`` int16 [N,2] (x,y),
``events_p.npy`` uint8 [N] in {0,1}, ``images.npy`` uint8 [F,H,W,1],
``images_ts.npy`` f64 [F,1], ``image_event_indices.npy`` int64 [F,1],
``metadata.json`` {"sensor_resolution": [H, W]}.

There is no network in the build/bench environment, so BASELINE.json's
configs are synthetic streams of the named shapes (uniform-random event
coordinates = worst case for the voxelizer's atomics; smooth moving texture for
the frames so SSIM denominators are well conditioned).
"""
import json
import os

import numpy as np


def make_stream(height, width, rate_ev_s, duration_s, fps, seed=0):
    """Returns dict of arrays in the on-disk layout."""
    g = np.random.default_rng(seed)
    n = int(rate_ev_s * duration_s)
    xs = g.integers(0, width, n, dtype=np.int64)
    ys = g.integers(0, height, n, dtype=np.int64)
    ts = np.sort(g.uniform(0.0, duration_s, n))
    ps = g.integers(0, 2, n, dtype=np.int64)
    num_frames = int(duration_s * fps)
    img_ts = (np.arange(num_frames, dtype=np.float64) + 1.0) / fps
    img_ts = img_ts[img_ts <= duration_s]
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    frames = np.empty((len(img_ts), height, width, 1), dtype=np.uint8)
    vx, vy = 40.0 + 10.0 * (seed % 5), 25.0 + 7.0 * (seed % 3)
    for i, t in enumerate(img_ts):
        tex = 0.5 + 0.4 * np.sin((xx + vx * t) / 9.0) * np.cos((yy + vy * t) / 7.0)
        frames[i, :, :, 0] = np.round(tex * 255.0).astype(np.uint8)
    idx = np.clip(np.searchsorted(ts, img_ts, 'right') - 1, 0, max(n - 1, 0))      # tools/bag_to_npy.py:80-81
    return {
        'events_ts': ts,
        'events_xy': np.stack([xs, ys], axis=1).astype(np.int16),
        'events_p': ps.astype(np.uint8),
        'images': frames,
        'images_ts': img_ts.reshape(-1, 1),
        'image_event_indices': idx.reshape(-1, 1).astype(np.int64),
        'sensor_resolution': [int(height), int(width)],
    }


def write_sequence(path, height, width, rate_ev_s, duration_s, fps, seed=0, with_images=True):
    os.makedirs(path, exist_ok=True)
    s = make_stream(height, width, rate_ev_s, duration_s, fps, seed)
    np.save(os.path.join(path, 'events_ts.npy'), s['events_ts'])
    np.save(os.path.join(path, 'events_xy.npy'), s['events_xy'])
    np.save(os.path.join(path, 'events_p.npy'), s['events_p'])
    if with_images:
        np.save(os.path.join(path, 'images.npy'), s['images'])
        np.save(os.path.join(path, 'images_ts.npy'), s['images_ts'])
        np.save(os.path.join(path, 'image_event_indices.npy'), s['image_event_indices'])
    with open(os.path.join(path, 'metadata.json'), 'w') as f:
        json.dump({'sensor_resolution': s['sensor_resolution']}, f)
    return path


# BASELINE.json configs: name -> (H, W, rate, duration, fps)
SHAPES = {
    'ecd': (180, 240, 1.0e6, 10.0, 24.0),        # cfg 2: E2VID on ECD-shape
    'hqf': (180, 240, 1.0e6, 5.0, 25.0),         # cfg 3: FireNet on 16 HQF-shape sequences (rate U(0.5,2) Mev/s)
    'mvsec': (260, 346, 5.0e6, 4.0, 45.0),       # cfg 4: HyperE2VID on MVSEC-shape
}


def hqf_rate(seed):
    """cfg 3: per-sequence rate ~ U(0.5, 2) Mev/s, seeded by the sequence index."""
    return float(np.random.default_rng(1000 + seed).uniform(0.5e6, 2.0e6))
