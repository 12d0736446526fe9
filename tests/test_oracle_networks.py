"""CPU: the functional network oracle against frames produced by the reference's real nn.Modules."""
import numpy as np
import torch

from helpers import golden, weights_of
from oracle import networks as on

torch.set_num_threads(4)


def _run(model, voxels):
    model.reset_states()
    return np.stack([model(torch.from_numpy(v)).numpy() for v in voxels])


def _check(tag, prefix, make, tol=2e-6, file='networks'):
    g = golden(file)
    _, w = weights_of(g, tag, prefix)
    got = _run(make(w), g[tag + '.voxels'])
    ref = g[tag + '.frames']
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= tol * max(1.0, np.max(np.abs(ref))), tag


def test_firenet_real_checkpoint():
    _check('firenet_ckpt', 'net.', lambda w: on.FireNetLegacyOracle(w))


def test_firenet_plus_real_checkpoint():
    _check('firenetplus_ckpt', '', lambda w: on.FireNetOracle(w))


def test_e2vid_topology():
    _check('e2vid_small', 'unetrecurrent.', lambda w: on.UNetRecurrentOracle(w, final_sigmoid=True))


def test_transposed_conv_decoder_topology():
    _check('e2vid_tconv', 'unetrecurrent.', lambda w: on.UNetRecurrentOracle(w, final_sigmoid=True, upsample_conv_decoder=False))


def test_e2vid_base32_one_encoder():
    """shipped width (last decoder 64 -> 32), one encoder: the fixture that pins the phase-stacked decoder (poly.cu)"""
    _check('e2vid_b32', 'unetrecurrent.', lambda w: on.UNetRecurrentOracle(w, 1, 1, final_sigmoid=True), file='networks_base32')
    _check('e2vid_b32_nonorm', 'unetrecurrent.', lambda w: on.UNetRecurrentOracle(w, 1, 1), file='networks_base32')


def test_flownet_topology():
    _check('flownet_small', 'unetflow.', lambda w: on.UNetRecurrentOracle(w))


def test_hyper_e2vid_topology():
    _check('hyper_small', 'unetrecurrent.', lambda w: on.UNetRecurrentOracle(w, dynamic_decoder=True))


def test_random_weight_builders_have_reference_shapes():
    w = on.random_unet_weights(seed=0, norm_bn=True)
    assert tuple(w['encoders.2.recurrent_block.Gates.weight'].shape) == (1024, 512, 3, 3)
    assert tuple(w['decoders.0.conv2d.weight'].shape) == (128, 256, 5, 5)
    assert 'encoders.0.conv.conv2d.bias' not in w and 'pred.norm_layer.running_var' in w
    w = on.random_unet_weights(seed=0, dynamic_decoder=True)
    assert tuple(w['decoders.0.dynamic_conv.compositional_coefficients'].shape) == (128, 1536, 1, 1)
    w = on.random_firenet_weights()
    assert tuple(w['head.recurrent_block.out_gate.weight'].shape) == (16, 32, 3, 3)


def test_real_e2vid_weights_two_encoder_slice():
    """REAL pretrained/E2VID weights (head, encoders 0-1, decoders 1-2, pred, eval-mode BatchNorm) through the real class"""
    _check('e2vid_real2', 'unetrecurrent.', lambda w: on.UNetRecurrentOracle(w, 2, 0, final_sigmoid=True), file='real_slices')


def test_real_e2vid_plus_weights_one_encoder_slice():
    _check('flownet_real1', 'unetflow.', lambda w: on.UNetRecurrentOracle(w, 1, 0), file='real_slices')


def test_spade_e2vid_oracle_vs_real_class():
    """SPADE-E2VID (model/spade_e2v.py) restatement against frames of the real Unet6 with seeded weights (regenerated here from
    the seed), three recurrent frames: the first takes the normalised-event branch of x_org, the others the previous output."""
    from evreal_b200 import synthetic
    g = golden('spade')
    o = on.SpadeE2vidOracle(synthetic.spade_state_dict(7))
    got = _run(o, g['seeded.voxels'])
    ref = g['seeded.frames']
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= 2e-6


def test_etnet_oracle_vs_real_class():
    """ET-Net (model/eitr/*) restatement against frames of the real EITR class with seeded weights (regenerated here from the
    seed): batch 2, 40x56 (35 tokens per scale), three recurrent frames."""
    from evreal_b200 import synthetic
    g = golden('etnet')
    o = on.ETNetOracle(synthetic.etnet_state_dict(9))
    got = _run(o, g['seeded.voxels'])
    ref = g['seeded.frames']
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= 2e-6
