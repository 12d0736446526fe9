import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REFERENCE = os.environ.get("EVREAL_REFERENCE", "/root/reference")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def gen_events(seed, n, H, W, dur=0.015):
    """SURVEY A.7 generator: draws in the order xs, ys, ts, ps."""
    g = np.random.default_rng(seed)
    xs = g.integers(0, W, n)
    ys = g.integers(0, H, n)
    ts = np.sort(g.uniform(0, dur, n))
    ps = g.integers(0, 2, n) * 2.0 - 1.0
    return (xs.astype(np.float32), ys.astype(np.float32), (ts - ts[0]).astype(np.float32), ps.astype(np.float32))


def weights_of(npz, tag, prefix):
    """{name-without-wrapper-prefix: tensor} for the oracle; {full name: tensor} is tag.w.<name>."""
    full = {k[len(tag) + 3:]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith(tag + ".w.")}
    stripped = {k[len(prefix):]: v for k, v in full.items()}
    return full, stripped


def write_sequence_from_arrays(path, arrays, sensor_resolution):
    import json
    os.makedirs(path, exist_ok=True)
    for k in ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices'):
        if k in arrays:
            np.save(os.path.join(path, k + '.npy'), arrays[k])
    with open(os.path.join(path, 'metadata.json'), 'w') as f:
        json.dump({'sensor_resolution': list(sensor_resolution)}, f)
    return path
