"""GPU parity with REAL pretrained weights at the shipped channel widths (round-1 verdict, weak #1: every full-width
test used seeded weights; the split-bf16 error budget depends on the operands' dynamic range).

  * real_slices.npz -- sub-networks made of real layers of pretrained/E2VID and pretrained/E2VID+, frames of the REAL
    reference classes over three recurrent steps (tools/make_golden.py::golden_real_slices); both precisions;
  * one real 256-channel residual convolution (+ folded BatchNorm) on a real bottleneck activation, through
    evk_conv2d_nhwc;
  * full_checkpoints.npz -- the complete shipped checkpoints (E2VID, E2VID+, HyperE2VID, SSL-E2VID, FireNet, FireNet+) at
    the BASELINE sizes through evaluate.get_model_from_checkpoint_path: the .pth files are not in git (43 MB each) but
    travel to the GPU box as untracked files under tests/golden/_ckpt (tools/make_golden.py copies them there); the test
    is skipped where they are absent.
Bar: max|diff| <= 1e-4 max|ref| per frame (north_star)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, gen_events, golden, weights_of

pytestmark = pytest.mark.gpu


def _frames(model, voxels):
    model.reset_states()
    return np.stack([model(torch.from_numpy(v).cuda())['image'].cpu().numpy() for v in voxels])


def _assert_close(got, ref, tol=1e-4):
    assert got.shape == ref.shape
    for f in range(ref.shape[0]):
        err = np.max(np.abs(got[f] - ref[f])) / max(np.max(np.abs(ref[f])), 1e-6)
        assert err <= tol, (f, err)


@pytest.mark.parametrize('precision', [0, 1])
def test_real_e2vid_slice(precision):
    from evreal_b200 import E2VIDRecurrent
    g = golden('real_slices')
    full, _ = weights_of(g, 'e2vid_real2', 'unetrecurrent.')
    kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 2, 'base_num_channels': 32,
          'num_residual_blocks': 0, 'use_upsample_conv': True, 'norm': 'BN', 'final_activation': 'sigmoid'}
    m = E2VIDRecurrent(kw).load_state_dict(full).to('cuda')
    m.precision = precision
    _assert_close(_frames(m, g['e2vid_real2.voxels']), g['e2vid_real2.frames'])
    if precision == 0:
        desc = ' | '.join(m.op_descriptions())
        assert 'stacked phases' in desc and 'tcgen05' in desc and 'lstm' in desc, desc


@pytest.mark.parametrize('precision', [0, 1])
def test_real_e2vid_plus_slice(precision):
    from evreal_b200 import FlowNet
    g = golden('real_slices')
    full, _ = weights_of(g, 'flownet_real1', 'unetflow.')
    base, nout = (int(v) for v in g['flownet_real1.kwargs'])
    kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 1, 'base_num_channels': base,
          'num_residual_blocks': 0, 'use_upsample_conv': True, 'norm': 'none', 'num_output_channels': nout}
    m = FlowNet(kw).load_state_dict(full).to('cuda')
    m.precision = precision
    _assert_close(_frames(m, g['flownet_real1.voxels']), g['flownet_real1.frames'])


@pytest.mark.parametrize('precision', [0, 1])
def test_real_residual_convolution_layer(precision):
    """resblocks.0.conv1 + bn1 + ReLU of pretrained/E2VID on the bottleneck activation the real network produced."""
    from evreal_b200 import _lib
    g = golden('real_slices')
    x = torch.from_numpy(g['resconv.x']).cuda().permute(0, 2, 3, 1).contiguous()          # NHWC
    w = g['resconv.weight'].astype(np.float64)
    scale = g['resconv.bn.weight'].astype(np.float64) / np.sqrt(g['resconv.bn.running_var'].astype(np.float64) + 1e-5)
    wf = np.ascontiguousarray((w * scale[:, None, None, None]).astype(np.float32))
    bf = np.ascontiguousarray((g['resconv.bn.bias'].astype(np.float64) - g['resconv.bn.running_mean'].astype(np.float64) * scale).astype(np.float32))
    N, H, W, C = x.shape
    y = torch.empty((N, H, W, wf.shape[0]), dtype=torch.float32, device='cuda')
    _lib.check(_lib.load().evk_conv2d_nhwc(_lib.ptr(x), N, H, W, C, wf.ctypes.data_as(ctypes.c_void_p), bf.ctypes.data_as(ctypes.c_void_p),
                                           wf.shape[0], 3, 1, 1, 1, None, precision, _lib.ptr(y), _lib.stream_ptr()))
    got = y.permute(0, 3, 1, 2).cpu().numpy()
    ref = g['resconv.y']
    assert np.max(np.abs(got - ref)) <= (3e-5 if precision == 0 else 2e-6) * np.max(np.abs(ref))


CKPT = os.path.join(GOLDEN, '_ckpt')


@pytest.mark.parametrize('name', ['E2VID', 'E2VID+', 'HyperE2VID', 'SSL-E2VID', 'FireNet', 'FireNet+'])
def test_full_shipped_checkpoint(name):
    path = os.path.join(CKPT, name + '.pth')
    if not os.path.exists(path):
        pytest.skip("shipped checkpoints are not in git; tools/make_golden.py --round2 copies them to tests/golden/_ckpt")
    from evreal_b200 import evaluate as ev
    from evreal_b200.util import CropParameters, normalize_pad
    from oracle import event_voxel as ov
    g = golden('full_checkpoints')
    H, W, norm_ev, n_enc = (int(v) for v in g[name + '.meta'])
    model = ev.get_model_from_checkpoint_path(name, path)
    assert model.num_encoders == n_enc
    crop = CropParameters(W, H, model.num_encoders)
    model.reset_states()
    ref = g[name + '.frames']
    for f in range(ref.shape[0]):
        e = gen_events(40 + f, 15000 + 7000 * f, H, W)
        v = ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in e], 5, (H, W))
        sums = g[name + '.voxel_sums'][f]
        assert abs(float(v.sum(dtype=torch.float64)) - sums[0]) <= 1e-6 * max(1.0, abs(sums[0])) + 1e-3
        assert abs(float(v.abs().sum(dtype=torch.float64)) - sums[1]) <= 1e-6 * sums[1]
        x = normalize_pad(v[None].cuda(), crop.height_crop_size, crop.width_crop_size, bool(norm_ev))
        got = crop.crop(model(x)['image'])[0, 0].cpu().numpy()
        err = np.max(np.abs(got - ref[f])) / np.max(np.abs(ref[f]))
        assert err <= 1e-4, (name, f, err)
