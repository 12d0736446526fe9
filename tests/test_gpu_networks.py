"""GPU parity: the CUDA network runner (evk_model_* through the host mirror classes) against
 (a) frames produced by the reference's real nn.Modules (tests/golden/networks.npz: real FireNet / FireNet+
     checkpoints, reduced-width E2VID / E2VID+ / HyperE2VID instances), and
 (b) the CPU oracle at the full BASELINE sizes with seeded random weights of the shipped shapes.

Tolerance (north_star): reconstructed frames within 1e-4 relative of the reference's fp32 torch path:
    max|diff| <= 1e-4 * max|ref|   per frame, over several recurrent steps.
"""
import numpy as np
import pytest
import torch

from helpers import golden, weights_of

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4


def _frames(model, voxels):
    model.reset_states()
    out = []
    for v in voxels:
        out.append(model(torch.from_numpy(np.ascontiguousarray(v)).cuda())['image'].cpu().numpy())
    return np.stack(out)


def _assert_close(got, ref, tag, tol=REL_TOL):
    assert got.shape == ref.shape, tag
    for f in range(ref.shape[0]):
        err = np.max(np.abs(got[f] - ref[f]))
        assert err <= tol * max(np.max(np.abs(ref[f])), 1e-3), (tag, f, err, np.max(np.abs(ref[f])))


def _load(model, full_weights):
    model.load_state_dict(full_weights)
    model.to('cuda')
    return model.eval()


E2VID_KW = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
            'num_residual_blocks': 2, 'use_upsample_conv': True}


@pytest.mark.parametrize('precision', [0, 1])
def test_firenet_real_checkpoint(precision):
    from evreal_b200 import FireNet_legacy
    g = golden('networks')
    full, _ = weights_of(g, 'firenet_ckpt', 'net.')
    m = _load(FireNet_legacy({'num_bins': 5, 'base_num_channels': 16, 'kernel_size': 3, 'recurrent_block_type': 'convgru',
                              'num_residual_blocks': 2, 'recurrent_blocks': {'resblock': [0]}}), full)
    m.precision = precision
    assert m.num_encoders == 4
    _assert_close(_frames(m, g['firenet_ckpt.voxels']), g['firenet_ckpt.frames'], 'firenet_ckpt')


@pytest.mark.parametrize('precision', [0, 1])
def test_firenet_plus_real_checkpoint_batch2(precision):
    from evreal_b200 import FireNet
    g = golden('networks')
    full, _ = weights_of(g, 'firenetplus_ckpt', '')
    m = _load(FireNet(), full)
    m.precision = precision
    _assert_close(_frames(m, g['firenetplus_ckpt.voxels']), g['firenetplus_ckpt.frames'], 'firenetplus_ckpt')


@pytest.mark.parametrize('precision', [0, 1])
def test_e2vid_topology_bn_sigmoid_batch2(precision):
    from evreal_b200 import E2VIDRecurrent
    g = golden('networks')
    full, _ = weights_of(g, 'e2vid_small', 'unetrecurrent.')
    m = _load(E2VIDRecurrent(dict(E2VID_KW, base_num_channels=8, norm='BN', final_activation='sigmoid')), full)
    m.precision = precision
    _assert_close(_frames(m, g['e2vid_small.voxels']), g['e2vid_small.frames'], 'e2vid_small')


@pytest.mark.parametrize('precision', [0, 1])
def test_transposed_conv_decoders(precision):
    """use_upsample_conv=False: TransposedConvLayer decoders (model/submodules.py:38-66) as zero insertion + flipped
    stride-1 convolution; frames of the real reference class with BN, batch 2, 3 recurrent steps."""
    from evreal_b200 import E2VIDRecurrent
    g = golden('networks')
    full, _ = weights_of(g, 'e2vid_tconv', 'unetrecurrent.')
    m = _load(E2VIDRecurrent(dict(E2VID_KW, use_upsample_conv=False, base_num_channels=8, norm='BN', final_activation='sigmoid')), full)
    m.precision = precision
    _assert_close(_frames(m, g['e2vid_tconv.voxels']), g['e2vid_tconv.frames'], 'e2vid_tconv')


@pytest.mark.parametrize('poly', [True, False])
@pytest.mark.parametrize('tag,norm,act', [('e2vid_b32', 'BN', 'sigmoid'), ('e2vid_b32_nonorm', 'none', '')])
def test_e2vid_base32_phase_stacked_decoder(tag, norm, act, poly, monkeypatch):
    """Shipped width (last decoder 64 -> 32) through the real reference class, one encoder, sizes that are not tile
    multiples (30x44, 18x26), batch 2, 3 recurrent frames: the last decoder runs as four stacked phases on the
    low-resolution map plus the border-ring correction (poly.cu); EVK_NO_POLY=1 is the upsample + 5x5 form."""
    from evreal_b200 import E2VIDRecurrent
    if not poly:
        monkeypatch.setenv('EVK_NO_POLY', '1')
    g = golden('networks_base32')
    full, _ = weights_of(g, tag, 'unetrecurrent.')
    m = _load(E2VIDRecurrent(dict(E2VID_KW, num_encoders=1, num_residual_blocks=1, base_num_channels=32, norm=norm,
                                  final_activation=act)), full)
    _assert_close(_frames(m, g[tag + '.voxels']), g[tag + '.frames'], tag)
    descs = m.op_descriptions()
    assert any('stacked phases' in d for d in descs) == poly, descs


def test_flownet_topology_image_channel():
    from evreal_b200 import FlowNet
    g = golden('networks')
    full, _ = weights_of(g, 'flownet_small', 'unetflow.')
    m = _load(FlowNet(dict(E2VID_KW, base_num_channels=4, norm='none', num_output_channels=3)), full)
    _assert_close(_frames(m, g['flownet_small.voxels']), g['flownet_small.frames'], 'flownet_small')


@pytest.mark.parametrize('precision', [0, 1])
def test_hyper_e2vid_topology_batch2(precision):
    from evreal_b200 import E2VIDRecurrent
    g = golden('networks')
    full, _ = weights_of(g, 'hyper_small', 'unetrecurrent.')
    m = _load(E2VIDRecurrent(dict(E2VID_KW, base_num_channels=4, norm='none', kernel_size=5, channel_multiplier=2,
                                  num_output_channels=1, use_dynamic_decoder=True)), full)
    m.precision = precision
    _assert_close(_frames(m, g['hyper_small.voxels']), g['hyper_small.frames'], 'hyper_small')


def _voxels(seed, frames, N, H, W, n_ev):
    from helpers import gen_events
    from oracle import event_voxel as ov
    out = []
    for f in range(frames):
        b = [ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in gen_events(seed + 31 * f + b_, n_ev, H, W)], 5, (H, W))
             for b_ in range(N)]
        out.append(torch.stack(b).numpy())
    return out


@pytest.mark.parametrize('norm_bn', [True, False])
def test_e2vid_two_encoders_phase_pair_decoder_vs_oracle(norm_bn):
    """Two encoders at the shipped width: decoders 128 -> 64 (one row phase per N tile, 4 of 5 tap rows) and 64 -> 32
    (four phases in one tile) both in phase-stacked form, sizes that are not tile multiples (36x52), batch 2, 3 frames."""
    from evreal_b200 import E2VIDRecurrent
    from oracle import networks as on
    w = on.random_unet_weights(seed=9, num_encoders=2, num_res=1, norm_bn=norm_bn)
    m = _load(E2VIDRecurrent(dict(E2VID_KW, num_encoders=2, num_residual_blocks=1, base_num_channels=32,
                                  norm='BN' if norm_bn else 'none', final_activation='sigmoid' if norm_bn else '')),
              {'unetrecurrent.' + k: v for k, v in w.items()})
    vox = _voxels(90, 3, 2, 36, 52, 900)
    oracle = on.UNetRecurrentOracle(w, 2, 1, final_sigmoid=norm_bn)
    ref = np.stack([oracle(torch.from_numpy(v)).numpy() for v in vox])
    _assert_close(_frames(m, vox), ref, 'e2vid_two_encoders')
    descs = m.op_descriptions()
    assert sum('stacked phases' in d for d in descs) == 2 and sum('tap rows per tile' in d for d in descs) == 1, descs


def test_e2vid_three_encoders_all_decoders_phase_stacked_vs_oracle():
    """Shipped E2VID shape at a small size that is not a tile multiple (40x56, batch 2, 3 frames, no norm): decoder 256 -> 128
    runs one output phase per N tile (4x4 of the 5x5 composite taps), 128 -> 64 one row phase per tile, 64 -> 32 all four."""
    from evreal_b200 import E2VIDRecurrent
    from oracle import networks as on
    w = on.random_unet_weights(seed=11, num_encoders=3, num_res=2, norm_bn=False)
    m = _load(E2VIDRecurrent(dict(E2VID_KW, base_num_channels=32, norm='none', final_activation='')),
              {'unetrecurrent.' + k: v for k, v in w.items()})
    vox = _voxels(95, 3, 2, 40, 56, 1200)
    oracle = on.UNetRecurrentOracle(w)
    ref = np.stack([oracle(torch.from_numpy(v)).numpy() for v in vox])
    _assert_close(_frames(m, vox), ref, 'e2vid_three_encoders')
    descs = m.op_descriptions()
    assert sum('stacked phases' in d for d in descs) == 3 and sum('4x4 of 5x5' in d for d in descs) == 1, descs


@pytest.mark.parametrize('precision', [0, 1])
def test_e2vid_full_size_vs_oracle(precision):
    """BASELINE cfg 2 shape: E2VID (BN, sigmoid, base 32) at 240x180 padded to 184x240, 5 recurrent frames."""
    from evreal_b200 import E2VIDRecurrent
    from oracle import networks as on
    w = on.random_unet_weights(seed=3, norm_bn=True)
    m = _load(E2VIDRecurrent(dict(E2VID_KW, base_num_channels=32, norm='BN', final_activation='sigmoid')),
              {'unetrecurrent.' + k: v for k, v in w.items()})
    m.precision = precision
    vox = _voxels(50, 5, 1, 184, 240, 40000)
    torch.set_num_threads(8)
    oracle = on.UNetRecurrentOracle(w, final_sigmoid=True)
    ref = np.stack([oracle(torch.from_numpy(v)).numpy() for v in vox])
    _assert_close(_frames(m, vox), ref, 'e2vid_full')
    assert m.flops_per_forward() == pytest.approx(40.10e9, rel=0.01)
    assert m.last_launch_count() > 0


def test_e2vid_and_firenet_at_the_largest_sensor_size_vs_oracle():
    """The largest size SURVEY 8a lists (640x480, cfg 5's sensor: 279.0 GFLOP per E2VID frame): two recurrent frames of E2VID
    and of FireNet against the CPU oracle -- every tile count, halo box and border line differs from the 240x180 runs."""
    from evreal_b200 import E2VIDRecurrent, FireNet_legacy
    from oracle import networks as on
    torch.set_num_threads(8)
    w = on.random_unet_weights(seed=13, norm_bn=True)
    m = _load(E2VIDRecurrent(dict(E2VID_KW, base_num_channels=32, norm='BN', final_activation='sigmoid')),
              {'unetrecurrent.' + k: v for k, v in w.items()})
    vox = _voxels(51, 2, 1, 480, 640, 300000)
    oracle = on.UNetRecurrentOracle(w, final_sigmoid=True)
    ref = np.stack([oracle(torch.from_numpy(v)).numpy() for v in vox])
    _assert_close(_frames(m, vox), ref, 'e2vid_vga')
    assert m.flops_per_forward() == pytest.approx(279.0e9, rel=0.01)
    wf = on.random_firenet_weights(seed=14)
    mf = _load(FireNet_legacy({'num_bins': 5, 'base_num_channels': 16, 'kernel_size': 3}), {'net.' + k: v for k, v in wf.items()})
    oracle_f = on.FireNetLegacyOracle(wf)
    ref = np.stack([oracle_f(torch.from_numpy(v)).numpy() for v in vox])
    _assert_close(_frames(mf, vox), ref, 'firenet_vga')


def test_firenet_full_size_vs_oracle_batch3():
    """BASELINE cfg 3 shape: FireNet at 240x180 padded to 192x240 (num_encoders falls back to 4), batch of 3 streams."""
    from evreal_b200 import FireNet_legacy
    from oracle import networks as on
    w = on.random_firenet_weights(seed=4)
    m = _load(FireNet_legacy({'num_bins': 5, 'base_num_channels': 16, 'kernel_size': 3}),
              {'net.' + k: v for k, v in w.items()})
    vox = _voxels(70, 4, 3, 192, 240, 30000)
    oracle = on.FireNetLegacyOracle(w)
    ref = np.stack([oracle(torch.from_numpy(v)).numpy() for v in vox])
    _assert_close(_frames(m, vox), ref, 'firenet_full')
    assert m.flops_per_forward() == pytest.approx(3 * 3.47e9, rel=0.01)


@pytest.mark.parametrize('precision', [0, 1])
def test_hyper_e2vid_full_size_vs_oracle(precision):
    """BASELINE cfg 4 shape: HyperE2VID at 346x260 padded to 264x352, 3 recurrent frames (prev_recs feeds frame 2+)."""
    from evreal_b200 import E2VIDRecurrent
    from oracle import networks as on
    w = on.random_unet_weights(seed=5, dynamic_decoder=True)
    m = _load(E2VIDRecurrent(dict(E2VID_KW, base_num_channels=32, norm='none', kernel_size=5, channel_multiplier=2,
                                  num_output_channels=1, use_dynamic_decoder=True)),
              {'unetrecurrent.' + k: v for k, v in w.items()})
    m.precision = precision
    vox = _voxels(90, 3, 1, 264, 352, 100000)
    torch.set_num_threads(8)
    oracle = on.UNetRecurrentOracle(w, dynamic_decoder=True)
    ref = np.stack([oracle(torch.from_numpy(v)).numpy() for v in vox])
    _assert_close(_frames(m, vox), ref, 'hyper_full')


def test_reset_states_and_state_roundtrip():
    from evreal_b200 import E2VIDRecurrent
    g = golden('networks')
    full, _ = weights_of(g, 'e2vid_small', 'unetrecurrent.')
    m = _load(E2VIDRecurrent(dict(E2VID_KW, base_num_channels=8, norm='BN', final_activation='sigmoid')), full)
    vox = g['e2vid_small.voxels']
    a = _frames(m, vox)
    b = _frames(m, vox)                       # reset_states() in between -> identical
    assert np.array_equal(a, b)
    # states property: run 2 frames, snapshot, run frame 3, restore, run frame 3 again
    m.reset_states()
    for v in vox[:2]:
        m(torch.from_numpy(v).cuda())
    snap = [(h.clone(), c.clone()) for h, c in m.states]
    assert tuple(snap[0][0].shape) == (2, 16, 16, 24) and tuple(snap[2][1].shape) == (2, 64, 4, 6)
    y1 = m(torch.from_numpy(vox[2]).cuda())['image'].clone()
    m.states = snap
    y2 = m(torch.from_numpy(vox[2]).cuda())['image']
    assert torch.equal(y1, y2)
    assert m.prev_recs is not None and tuple(m.prev_recs.shape) == (2, 1, 32, 48)


def test_firenet_state_roundtrip_in_window_mode():
    """FireNet's ConvGRU states through the `states` property while the 16-channel layers run in window mode (row-padded
    split companions: evk_model_set_state re-splits into the padded layout): snapshot after 2 frames, run frame 3,
    restore, run frame 3 again -> bit-identical; reset_states() in between -> identical sequences."""
    from evreal_b200 import FireNet_legacy
    g = golden('networks')
    full, _ = weights_of(g, 'firenet_ckpt', 'net.')
    m = _load(FireNet_legacy({'num_bins': 5, 'base_num_channels': 16, 'kernel_size': 3, 'recurrent_block_type': 'convgru',
                              'num_residual_blocks': 2, 'recurrent_blocks': {'resblock': [0]}}), full)
    vox = g['firenet_ckpt.voxels']
    a = _frames(m, vox)
    assert np.array_equal(a, _frames(m, vox))
    assert any('window K rows' in d for d in m.op_descriptions())
    m.reset_states()
    for v in vox[:2]:
        m(torch.from_numpy(np.ascontiguousarray(v)).cuda())
    snap = [s.clone() for s in m.states]
    assert len(snap) == 2 and tuple(snap[0].shape) == (1, 16, 48, 64)
    y1 = m(torch.from_numpy(np.ascontiguousarray(vox[2])).cuda())['image'].clone()
    m.states = snap
    y2 = m(torch.from_numpy(np.ascontiguousarray(vox[2])).cuda())['image']
    assert torch.equal(y1, y2)
    assert np.array_equal(y1.cpu().numpy(), a[2])


def test_batching_is_parity_safe():
    """two streams as one batch == the same streams run separately."""
    from evreal_b200 import E2VIDRecurrent
    g = golden('networks')
    full, _ = weights_of(g, 'e2vid_small', 'unetrecurrent.')
    vox = g['e2vid_small.voxels']
    kw = dict(E2VID_KW, base_num_channels=8, norm='BN', final_activation='sigmoid')
    both = _frames(_load(E2VIDRecurrent(kw), full), vox)
    for b in range(2):
        single = _frames(_load(E2VIDRecurrent(kw), full), vox[:, b:b + 1])
        assert np.max(np.abs(single[:, 0] - both[:, b])) <= 1e-6


def test_loader_errors():
    from evreal_b200 import E2VIDRecurrent, _lib
    g = golden('networks')
    full, _ = weights_of(g, 'e2vid_small', 'unetrecurrent.')
    kw = dict(E2VID_KW, base_num_channels=8, norm='BN', final_activation='sigmoid')
    broken = {k: v for k, v in full.items() if 'encoders.1.recurrent_block.Gates.weight' not in k}
    m = _load(E2VIDRecurrent(kw), broken)
    with pytest.raises(KeyError):
        m(torch.zeros(1, 5, 32, 48).cuda())
    m = _load(E2VIDRecurrent(kw), full)
    with pytest.raises(ValueError):                      # not a multiple of 2^num_encoders: CropParameters.pad was skipped
        m(torch.zeros(1, 5, 30, 48).cuda())
    with pytest.raises(RuntimeError):
        E2VIDRecurrent(kw).load_state_dict({'bogus.weight': torch.zeros(1)})
    with pytest.raises(_lib.EvkError):
        E2VIDRecurrent(kw).to('cpu')
