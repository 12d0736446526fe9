"""GPU parity of ET-Net (model/eitr/*: EITR over mls_tpa): E2VID's head / recurrent encoders / upsample-conv decoders
around a three-scale transformer token path (pre-norm encoder layers, decoders with cross attention to the coarser scale).
Against frames of the REAL class (tests/golden/etnet.npz, tools/make_golden.py::golden_etnet): seeded weights regenerated from
the seed, and the shipped checkpoint at 180x240 (690 tokens per scale) when its .pth travelled to the box (tests/golden/_ckpt,
untracked)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, gen_events, golden

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=1e-4):
    assert got.shape == ref.shape
    err = np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-6)
    assert err <= tol, err


def test_etnet_seeded_weights_vs_real_class():
    from evreal_b200 import EITR, synthetic
    g = golden('etnet')
    m = EITR({'num_bins': 5, 'norm': None}).load_state_dict(synthetic.etnet_state_dict(9)).to('cuda')
    m.reset_states()
    for v, want in zip(g['seeded.voxels'], g['seeded.frames']):
        _close(m(torch.from_numpy(v).cuda())['image'].cpu().numpy(), want)
    desc = ' | '.join(m.op_descriptions())
    assert 'attention 8 heads' in desc and 'LayerNorm(256)' in desc and 'lstm' in desc and 'tcgen05' in desc, desc
    st = m.states
    assert len(st) == 3 and tuple(st[0][0].shape) == (2, 64, 20, 28) and tuple(st[2][1].shape) == (2, 256, 5, 7)
    # a second sequence starts from zero state again
    m.reset_states()
    _close(m(torch.from_numpy(g['seeded.voxels'][0]).cuda())['image'].cpu().numpy(), g['seeded.frames'][0])


def test_etnet_fp32_path_and_state_round_trip():
    """precision = 1 (every convolution / linear layer on the exact-fp32 CUDA-core kernel) agrees too, and states set on a fresh
    program continue the sequence."""
    from evreal_b200 import EITR, synthetic
    g = golden('etnet')
    sd = synthetic.etnet_state_dict(9)
    m = EITR({'num_bins': 5, 'norm': None}).load_state_dict(sd).to('cuda')
    m.precision = 1
    m.reset_states()
    vox = [torch.from_numpy(v).cuda() for v in g['seeded.voxels']]
    _close(m(vox[0])['image'].cpu().numpy(), g['seeded.frames'][0], 2e-5)
    _close(m(vox[1])['image'].cpu().numpy(), g['seeded.frames'][1], 2e-5)
    st = m.states
    m2 = EITR({'num_bins': 5, 'norm': None}).load_state_dict(sd).to('cuda')
    m2(vox[0])                                   # fixes the shapes
    m2.states = st
    _close(m2(vox[2])['image'].cpu().numpy(), g['seeded.frames'][2])


def test_etnet_shipped_checkpoint_full_size():
    path = os.path.join(GOLDEN, '_ckpt', 'ET-Net.pth')
    if not os.path.exists(path):
        pytest.skip("the shipped checkpoint is not in git; tools/make_golden.py --only-etnet copies it to tests/golden/_ckpt")
    from evreal_b200 import evaluate as ev
    from evreal_b200.util import CropParameters, normalize_pad
    from oracle import event_voxel as ov
    g = golden('etnet')
    H, W = 180, 240
    model = ev.get_model_from_checkpoint_path('ET-Net', path)
    assert model.num_encoders == 3
    crop = CropParameters(W, H, 3)
    model.reset_states()
    for f, want in enumerate(g['ckpt.frames']):
        e = gen_events(40 + f, 15000 + 7000 * f, H, W)
        v = ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in e], 5, (H, W))
        assert abs(float(v.abs().sum(dtype=torch.float64)) - g['ckpt.voxel_sums'][f][1]) <= 1e-6 * g['ckpt.voxel_sums'][f][1]
        x = normalize_pad(v[None].cuda(), crop.height_crop_size, crop.width_crop_size, False)
        _close(crop.crop(model(x)['image'])[0, 0].cpu().numpy(), want)
