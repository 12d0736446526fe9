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


# ---------------------------------------------------------------------------
# Seeded random-initialised weights with the shipped checkpoints' names and shapes.  There is no
# network for the real checkpoints on the GPU box, so bench.py times these architectures with these weights
# (same FLOPs, same kernels); the dicts load into the host mirror classes through load_state_dict.
# ---------------------------------------------------------------------------
def _conv_w(g, out_c, in_c, k, bias=True, gain=1.0):
    import torch
    d = {'weight': torch.from_numpy((g.standard_normal((out_c, in_c, k, k)) * (gain / (in_c * k * k) ** 0.5)).astype(np.float32))}
    if bias:
        d['bias'] = torch.from_numpy((g.standard_normal(out_c) * 0.1).astype(np.float32))
    return d


def _bn_w(g, c):
    import torch
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    return {'weight': f(1.0 + 0.2 * g.standard_normal(c)), 'bias': f(0.1 * g.standard_normal(c)),
            'running_mean': f(0.1 * g.standard_normal(c)), 'running_var': f(0.5 + g.random(c))}


def unet_state_dict(seed=0, prefix='unetrecurrent.', base=32, num_encoders=3, num_res=2, k=5, bins=5, norm_bn=False,
                    num_out=1, dynamic_decoder=False):
    """E2VID (norm_bn=True) / E2VID+ / SSL-E2VID / HyperE2VID (dynamic_decoder=True) state_dict."""
    import torch
    g = np.random.default_rng(seed)
    w = {}

    def put(pfx, d):
        for kk, v in d.items():
            w[prefix + pfx + '.' + kk] = v
    put('head.conv2d', _conv_w(g, base, bins, k))
    cin = base
    for i in range(num_encoders):
        cout = cin * 2
        put('encoders.%d.conv.conv2d' % i, _conv_w(g, cout, cin, k, bias=not norm_bn, gain=1.4))
        if norm_bn:
            put('encoders.%d.conv.norm_layer' % i, _bn_w(g, cout))
        put('encoders.%d.recurrent_block.Gates' % i, _conv_w(g, 4 * cout, 2 * cout, 3))
        cin = cout
    for j in range(num_res):
        for n, b in (('conv1', 'bn1'), ('conv2', 'bn2')):
            put('resblocks.%d.%s' % (j, n), _conv_w(g, cin, cin, 3, bias=not norm_bn))
            if norm_bn:
                put('resblocks.%d.%s' % (j, b), _bn_w(g, cin))
    for i in range(num_encoders):
        cout = cin // 2
        if i == 0 and dynamic_decoder:
            p = 'decoders.0'
            put(p + '.context_fusion.conv', _conv_w(g, 32, bins + 1, 3))
            put(p + '.dynamic_atom_generation.bases_net.0', _conv_w(g, 64, 32, 3))
            put(p + '.dynamic_atom_generation.bases_net.1', _bn_w(g, 64))
            put(p + '.dynamic_atom_generation.bases_net.3', _conv_w(g, 72, 64, 3))
            put(p + '.dynamic_atom_generation.bases_net.4', _bn_w(g, 72))
            w[prefix + p + '.dynamic_atom_generation.bases'] = torch.from_numpy((g.standard_normal((12, k * k)) * 0.3).astype(np.float32))
            w[prefix + p + '.dynamic_conv.compositional_coefficients'] = torch.from_numpy(
                (g.standard_normal((cout, cin * 6, 1, 1)) * (1.4 / (cin * 6) ** 0.5)).astype(np.float32))
            w[prefix + p + '.dynamic_conv.bias'] = torch.from_numpy((g.standard_normal(cout) * 0.1).astype(np.float32))
        else:
            put('decoders.%d.conv2d' % i, _conv_w(g, cout, cin, k, bias=not norm_bn, gain=1.4))
            if norm_bn:
                put('decoders.%d.norm_layer' % i, _bn_w(g, cout))
        cin = cout
    put('pred.conv2d', _conv_w(g, num_out, base, 1, bias=not norm_bn))
    if norm_bn:
        put('pred.norm_layer', _bn_w(g, num_out))
    return w


def firenet_state_dict(seed=0, prefix='net.', base=16, bins=5):
    """pretrained/FireNet (FireNet_legacy) state_dict."""
    g = np.random.default_rng(seed)
    w = {}

    def put(pfx, d):
        for kk, v in d.items():
            w[prefix + pfx + '.' + kk] = v
    put('head.conv.conv2d', _conv_w(g, base, bins, 3))
    for unit in ('head.recurrent_block', 'resblocks.0.recurrent_block'):
        for gate in ('reset_gate', 'update_gate', 'out_gate'):
            put(unit + '.' + gate, _conv_w(g, base, 2 * base, 3, gain=1.4))
    for blk in ('resblocks.0.conv', 'resblocks.1'):
        for c in ('conv1', 'conv2'):
            put(blk + '.' + c, _conv_w(g, base, base, 3, gain=1.2))
    put('pred.conv2d', _conv_w(g, 1, base, 1))
    return w


def spade_state_dict(seed=0):
    """pretrained/SPADE-E2VID (model/spade_e2v.py Unet6) state_dict: the class has no width parameters."""
    import torch
    g = np.random.default_rng(seed)
    w = {}

    def put(pfx, d):
        for kk, v in d.items():
            w[pfx + '.' + kk] = v

    def bn(c, affine=True):
        d = _bn_w(g, c)
        if not affine:
            d = {k: v for k, v in d.items() if k.startswith('running')}
        return d
    put('fc', _conv_w(g, 32, 5, 5))
    for name, cin, cout in (('rec0', 32, 64), ('rec1', 64, 128), ('rec2', 128, 256), ('up2', 64, 32)):
        put(name + '.conv0', _conv_w(g, cout, cin, 5, bias=False, gain=1.4))
        put(name + '.bn', bn(cout))
        put(name + '.recurrent_block.Gates', _conv_w(g, 4 * cout, 2 * cout, 3))
    for name in ('res0', 'res1'):
        for c, b in (('conv1', 'bn1'), ('conv2', 'bn2')):
            put(name + '.' + c, _conv_w(g, 256, 256, 3, bias=False))
            put(name + '.' + b, bn(256))
    for name, cin, cout in (('up0', 256, 128), ('up1', 128, 64)):
        put(name + '.conv0', _conv_w(g, 4 * cout, cin, 3, bias=False, gain=1.4))
        put(name + '.norm.param_free_norm', bn(cout, affine=False))
        put(name + '.norm.mlp_shared.0', _conv_w(g, 64, 3, 3))
        put(name + '.norm.mlp_gamma', _conv_w(g, cout, 64, 3, gain=0.5))
        put(name + '.norm.mlp_beta', _conv_w(g, cout, 64, 3, gain=0.5))
    put('conv_img', _conv_w(g, 3, 32, 1))
    put('bn_img', bn(3))
    return w


def etnet_state_dict(seed=0):
    """pretrained/ET-Net (model/eitr/u_trans.py mls_tpa, norm None) state_dict: fixed widths."""
    import torch
    g = np.random.default_rng(seed)
    w = {}
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))

    def put(pfx, d):
        for kk, v in d.items():
            w[pfx + '.' + kk] = v

    def linear(pfx, out_c, in_c, gain=1.0):
        w[pfx + '.weight'] = f32(g.standard_normal((out_c, in_c)) * (gain / in_c ** 0.5))
        w[pfx + '.bias'] = f32(g.standard_normal(out_c) * 0.05)

    def ln(pfx):
        w[pfx + '.weight'] = f32(1.0 + 0.1 * g.standard_normal(256))
        w[pfx + '.bias'] = f32(0.05 * g.standard_normal(256))

    def mha(pfx):
        w[pfx + '.in_proj_weight'] = f32(g.standard_normal((768, 256)) / 16.0)
        w[pfx + '.in_proj_bias'] = f32(g.standard_normal(768) * 0.05)
        linear(pfx + '.out_proj', 256, 256)
    put('head.conv2d', _conv_w(g, 32, 5, 5))
    cin = 32
    for i in range(3):
        put('DownsampleConv.%d.conv.conv2d' % i, _conv_w(g, 2 * cin, cin, 5, gain=1.4))
        put('DownsampleConv.%d.recurrent_block.Gates' % i, _conv_w(g, 8 * cin, 4 * cin, 3))
        cin *= 2
    put('split1', _conv_w(g, 256, 128, 2))
    put('split2', _conv_w(g, 256, 64, 4))
    for s_ in range(3):
        for i in range(3):
            p = 'trans_encoder%d.encoder.layers.%d' % (s_, i)
            mha(p + '.self_attn'); ln(p + '.norm1'); ln(p + '.norm2')
            linear(p + '.linear1', 1024, 256, 1.4); linear(p + '.linear2', 256, 1024)
        for i in range(2):
            p = 'trans_decoder%d.decoder.layers.%d' % (s_, i)
            mha(p + '.self_attn'); mha(p + '.cross_attn')
            for nn_ in ('norm1', 'norm21', 'norm22', 'norm3'):
                ln(p + '.' + nn_)
            linear(p + '.linear1', 1024, 256, 1.4); linear(p + '.linear2', 256, 1024)
    for i in range(3):
        put('UpsampleConv.%d.conv2d' % i, _conv_w(g, cin // 2, cin, 5, gain=1.4))
        cin //= 2
    put('pred.conv2d', _conv_w(g, 1, 32, 1))
    return w


E2VID_KWARGS = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
                'base_num_channels': 32, 'num_residual_blocks': 2, 'use_upsample_conv': True, 'norm': 'BN',
                'final_activation': 'sigmoid'}
HYPER_KWARGS = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'kernel_size': 5,
                'channel_multiplier': 2, 'num_encoders': 3, 'base_num_channels': 32, 'num_residual_blocks': 2,
                'use_upsample_conv': True, 'norm': 'none', 'num_output_channels': 1, 'use_dynamic_decoder': True}
FIRENET_KWARGS = {'num_bins': 5, 'base_num_channels': 16, 'kernel_size': 3, 'recurrent_block_type': 'convgru',
                  'num_residual_blocks': 2, 'recurrent_blocks': {'resblock': [0]}}
