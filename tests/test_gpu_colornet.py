"""GPU: ColorNet (model/model.py:46-105) -- the Bayer split of a CED-style event tensor reconstructed as one
batch-4 forward (R, G, B, W sites) plus one batch-1 forward (grey) and merged on the host, against frames of the REAL
ColorNet wrapped around the real FireNet checkpoint (tests/golden/colornet.npz, tools/make_golden.py::golden_colornet).
The reference quantises every site image with astype(uint8) (truncation) before merging, so a 1e-6 difference in a
reconstruction can move single pixels by one grey level: the bar is 'almost all pixels identical, none off by more than
3 levels'."""
import numpy as np
import pytest
import torch

from helpers import golden, weights_of

pytestmark = pytest.mark.gpu


def test_colornet_matches_the_real_class():
    from evreal_b200 import FireNet_legacy
    from evreal_b200.model import ColorNet
    g = golden('colornet')
    full, _ = weights_of(golden('networks'), 'firenet_ckpt', 'net.')
    m = FireNet_legacy({'num_bins': 5, 'base_num_channels': 16, 'kernel_size': 3}).load_state_dict(full).to('cuda')
    cn = ColorNet(m)
    cn.reset_states()
    for v, want in zip(g['voxels'], g['frames']):
        got = cn(torch.from_numpy(v))['image'].numpy()
        assert got.shape == want.shape == (3, 40, 56)
        levels = np.abs(np.round(got * 255) - np.round(want * 255))
        assert levels.max() <= 3, levels.max()
        assert (levels == 0).mean() >= 0.99, (levels == 0).mean()
    # a second sequence: states of both device programs are cleared
    cn.reset_states()
    again = cn(torch.from_numpy(g['voxels'][0]))['image'].numpy()
    assert np.abs(np.round(again * 255) - np.round(g['frames'][0] * 255)).max() <= 3


def test_colour_sequence_through_evaluate(tmp_path):
    """eval config with color: true (config/eval/color.json): frames are written, no quantitative scores (utils/eval_metrics.py:262)"""
    import json
    import os
    from evreal_b200 import evaluate as ev
    from helpers import write_sequence_from_arrays
    g = golden('eval_loop')
    full, _ = weights_of(golden('networks'), 'firenet_ckpt', 'net.')
    arrays = {k: g[k] for k in ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices')}
    path = write_sequence_from_arrays(str(tmp_path / 'seq'), arrays, (48, 64))
    from evreal_b200 import FireNet_legacy
    from evreal_b200.model import ColorNet
    m = ColorNet(FireNet_legacy({'num_bins': 5, 'base_num_channels': 16, 'kernel_size': 3}).load_state_dict(full).to('cuda'))
    cfg = {'name': 'color', 'save_images': True, 'histeq': 'none', 'color': True, 'eval_infer_all': False, 'ts_tol_ms': 1.0,
           'create_video': False, 'dataset_kwargs': {'num_bins': 5, 'voxel_method': {'method': 'between_frames'}}}
    seq = {'name': 'seq', 'sequence_path': path, 'start_time_s': 0.2, 'end_time_s': 0.5, 'dataset_kwargs': dict(cfg['dataset_kwargs'])}
    n_eval, means, frames, _ = ev.eval_method_on_sequence('CED', cfg, 'FireNet', m, {'event_tensor_normalization': True}, seq,
                                                          ['mse', 'ssim'], output_root=str(tmp_path / 'out'), write_files=True)
    assert n_eval == 0 and frames > 0
    out = tmp_path / 'out' / 'color' / 'CED' / 'seq' / 'FireNet'
    pngs = sorted(f for f in os.listdir(out) if f.endswith('.png'))
    assert len(pngs) == frames
    import cv2
    assert cv2.imread(str(out / pngs[-1]), cv2.IMREAD_UNCHANGED).shape == (48, 64, 3)
