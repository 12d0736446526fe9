"""GPU parity of SPADE-E2VID (model/spade_e2v.py Unet6): stride-1 recurrent encoder at full resolution, pixel-shuffle
decoders with SPADE normalisation conditioned on the previous reconstruction, recurrent last decoder, 3-channel sigmoid head.
Against frames of the REAL class (tests/golden/spade.npz, tools/make_golden.py::golden_spade): seeded weights regenerated from
the seed, and the shipped checkpoint at 180x240 when its .pth travelled to the box (tests/golden/_ckpt, untracked)."""
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


def test_spade_seeded_weights_vs_real_class():
    from evreal_b200 import SpadeE2vid, synthetic
    g = golden('spade')
    m = SpadeE2vid().load_state_dict(synthetic.spade_state_dict(7)).to('cuda')
    m.reset_states()
    for v, want in zip(g['seeded.voxels'], g['seeded.frames']):
        _close(m(torch.from_numpy(v).cuda())['image'].cpu().numpy(), want)
    desc = ' | '.join(m.op_descriptions())
    assert 'pixel shuffle x2 + SPADE' in desc and 'lstm' in desc and 'tcgen05' in desc, desc
    # a second sequence starts from the normalised-event branch again
    m.reset_states()
    _close(m(torch.from_numpy(g['seeded.voxels'][0]).cuda())['image'].cpu().numpy(), g['seeded.frames'][0])
    # states round trip (model.states get / set: four (hidden, cell) pairs)
    st = m.states
    assert len(st) == 4 and tuple(st[0][0].shape) == (1, 64, 40, 56) and tuple(st[3][1].shape) == (1, 32, 40, 56)


def test_spade_batch_is_per_sample():
    """Two different streams in one batch: each gets the frames it gets alone (the first-frame min / max of x[:, :3] is per
    sample here; the reference computes it over the tensor and only ever runs batch 1)."""
    from evreal_b200 import SpadeE2vid, synthetic
    from oracle import networks as on
    sd = synthetic.spade_state_dict(7)
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(2, 5, 24, 40, generator=g) * torch.tensor([0.4, 1.3]).view(2, 1, 1, 1) for _ in range(3)]
    m = SpadeE2vid().load_state_dict(sd).to('cuda')
    m.reset_states()
    got = [m(x.cuda())['image'].cpu().numpy() for x in xs]
    for b in range(2):
        o = on.SpadeE2vidOracle(sd)
        for f, x in enumerate(xs):
            _close(got[f][b:b + 1], o(x[b:b + 1]).numpy())


def test_spade_shipped_checkpoint_full_size():
    path = os.path.join(GOLDEN, '_ckpt', 'SPADE-E2VID.pth')
    if not os.path.exists(path):
        pytest.skip("the shipped checkpoint is not in git; tools/make_golden.py --only-spade copies it to tests/golden/_ckpt")
    from evreal_b200 import evaluate as ev
    from evreal_b200.util import CropParameters, normalize_pad
    from oracle import event_voxel as ov
    g = golden('spade')
    H, W = 180, 240
    model = ev.get_model_from_checkpoint_path('SPADE-E2VID', path)
    assert model.num_encoders == 3
    crop = CropParameters(W, H, 3)
    model.reset_states()
    for f, want in enumerate(g['ckpt.frames']):
        e = gen_events(40 + f, 15000 + 7000 * f, H, W)
        v = ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in e], 5, (H, W))
        assert abs(float(v.abs().sum(dtype=torch.float64)) - g['ckpt.voxel_sums'][f][1]) <= 1e-6 * g['ckpt.voxel_sums'][f][1]
        x = normalize_pad(v[None].cuda(), crop.height_crop_size, crop.width_crop_size, False)
        _close(crop.crop(model(x)['image'])[0, 0].cpu().numpy(), want)
