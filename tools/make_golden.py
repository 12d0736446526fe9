"""Generates tests/golden/*.npz by running the REAL reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tools/make_golden.py
Every vector below comes out of the reference's own modules -- utils.event_utils,
dataset.MemMapDataset, utils.util.CropParameters, eval.normalize_event_tensor,
model.* (real classes; real FireNet checkpoint, reduced-width random-weight
instances of the E2VID family so the fixtures stay small) and the real
eval.eval_method_on_sequence loop driven through the import shims of SURVEY 8c
(yachalk / pyiqa / ffmpeg stubs, skimage.metrics restated with scipy-equivalent
code from oracle/metrics.py, CudaTimer -> nullcontext, weights_only=False).
"""
import contextlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EVREAL_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from evreal_b200 import synthetic  # noqa: E402  (pure numpy generator, no CUDA needed)
from oracle import metrics as om  # noqa: E402


def install_shims():
    class _Chalk:
        def __getattr__(self, _):
            return self

        def __call__(self, s):
            return str(s)

    sys.modules['yachalk'] = types.SimpleNamespace(chalk=_Chalk())
    sys.modules['pyiqa'] = types.SimpleNamespace(list_models=lambda: [])
    sys.modules['ffmpeg'] = types.ModuleType('ffmpeg')
    sk = types.ModuleType('skimage')
    skm = types.ModuleType('skimage.metrics')
    skm.mean_squared_error = lambda a, b: om.mse_oracle(b, a)
    skm.structural_similarity = lambda a, b, **kw: om.ssim_oracle(b, a, kw.get('data_range', 1.0))
    sk.metrics = skm
    sys.modules['skimage'] = sk
    sys.modules['skimage.metrics'] = skm
    _load = torch.load
    torch.load = lambda *a, **k: _load(*a, **{**k, 'weights_only': False})


def gen_events(seed, n, H, W, dur=0.015):
    g = np.random.default_rng(seed)
    xs = g.integers(0, W, n)
    ys = g.integers(0, H, n)
    ts = np.sort(g.uniform(0, dur, n))
    ps = g.integers(0, 2, n) * 2.0 - 1.0
    return (xs.astype(np.float32), ys.astype(np.float32), (ts - ts[0]).astype(np.float32), ps.astype(np.float32))


def golden_voxel():
    from utils.event_utils import events_to_voxel_torch
    cases = {}

    def add(name, xs, ys, ts, ps, H, W, bins=5):
        grid = events_to_voxel_torch(*[torch.from_numpy(np.asarray(v, dtype=np.float32)) for v in (xs, ys, ts, ps)],
                                     bins, sensor_size=(H, W)).numpy()
        for k, v in (('xs', xs), ('ys', ys), ('ts', ts), ('ps', ps), ('grid', grid)):
            cases[f'{name}.{k}'] = np.asarray(v, dtype=np.float32)
        cases[f'{name}.meta'] = np.array([H, W, bins], dtype=np.int64)

    add('small', *gen_events(1, 500, 24, 32), 24, 32)
    add('cfg1', *gen_events(0, 15000, 180, 240), 180, 240)            # SURVEY A.7 smoke case
    add('bins3', *gen_events(2, 700, 16, 20), 16, 20, bins=3)
    add('one_event', [3], [2], [0.0], [1.0], 8, 8)
    add('two_equal_t', [1, 2], [1, 2], [0.0, 0.0], [1.0, -1.0], 8, 8)  # dt < 1e-9 -> linspace
    add('three_equal_t', [1, 2, 3], [1, 2, 3], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], 8, 8)
    add('negative_wrap', [-1, 0, 5], [0, -2, 3], [0.0, 0.5, 1.0], [1.0, 1.0, -1.0], 8, 8)
    add('same_pixel', np.full(64, 3), np.full(64, 4), np.linspace(0, 1, 64), np.ones(64), 8, 8)
    np.savez_compressed(os.path.join(OUT, 'voxel.npz'), **cases)
    print('voxel.npz: cfg1 sum=%.4f abs=%.2f' % (cases['cfg1.grid'].sum(), np.abs(cases['cfg1.grid']).sum()))


def golden_glue():
    """normalize_event_tensor + CropParameters on a few shapes."""
    import eval as ref_eval
    from utils.util import CropParameters
    out = {}
    for name, (H, W, enc) in {'e2vid_180x240': (180, 240, 3), 'firenet_180x240': (180, 240, 4),
                              'mvsec_260x346': (260, 346, 3), 'odd_37x53': (37, 53, 2)}.items():
        cp = CropParameters(W, H, enc)
        out[f'{name}.meta'] = np.array([H, W, enc, cp.height_crop_size, cp.width_crop_size, cp.padding_top,
                                        cp.padding_left, cp.iy0, cp.iy1, cp.ix0, cp.ix1], dtype=np.int64)
    from utils.event_utils import events_to_voxel_torch
    v = events_to_voxel_torch(*[torch.from_numpy(a) for a in gen_events(3, 4000, 37, 53)], 5, sensor_size=(37, 53))
    vn = ref_eval.normalize_event_tensor(v[None])
    cp = CropParameters(53, 37, 2)
    out['norm.in'] = v.numpy()
    out['norm.out'] = vn.numpy()
    out['norm.padded'] = cp.pad(vn).numpy()
    out['norm.cropped_back'] = cp.crop(cp.pad(vn)).numpy()
    np.savez_compressed(os.path.join(OUT, 'glue.npz'), **out)


def golden_windows(tmp):
    from dataset import MemMapDataset
    path = synthetic.write_sequence(os.path.join(tmp, 'winseq'), 32, 40, 20000.0, 1.5, 20.0, seed=5)
    out = {k: np.load(os.path.join(path, k + '.npy')) for k in
           ('events_ts', 'events_xy', 'events_p', 'images_ts', 'image_event_indices')}
    modes = {
        'between_frames': {'method': 'between_frames'},
        'k_events': {'method': 'k_events', 'k': 1500, 'sliding_window_w': 0},
        'k_events_sliding': {'method': 'k_events', 'k': 1500, 'sliding_window_w': 500},
        't_seconds': {'method': 't_seconds', 't': 0.04, 'sliding_window_t': 0.0},
        't_seconds_sliding': {'method': 't_seconds', 't': 0.05, 'sliding_window_t': 0.01},
    }
    for name, vm in modes.items():
        ds = MemMapDataset(path, voxel_method=dict(vm), num_bins=5)
        rows = []
        for i in range(len(ds)):
            try:
                item = ds[i]
            except ValueError:
                rows.append([i, -1, -1, 0, 0, 0])       # window past the end of the stream (dataset.py:196-197)
                continue
            if vm['method'] == 'between_frames':
                prev = ds.frames_to_use[i - 1] if i > 0 else 0
                idx0, idx1 = ds.event_indices[prev][1], ds.event_indices[ds.frames_to_use[i]][1]
            else:
                idx0, idx1 = ds.event_indices[i]
            rows.append([i, idx0, idx1, item['event_count'], item['voxel_timestamp'].item(), item['dt'].item()])
        out[f'{name}.items'] = np.array(rows, dtype=np.float64)
        out[f'{name}.table'] = np.array(ds.event_indices, dtype=np.int64)
        out[f'{name}.len'] = np.array([len(ds)], dtype=np.int64)
        if vm['method'] != 'between_frames':
            out[f'{name}.closest_frame'] = np.array(
                [ds.get_closest_frame_index(r[4]) for r in rows if r[1] >= 0], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, 'windows.npz'), **out)


def _run_frames(model, voxels):
    outs = []
    model.reset_states()
    with torch.no_grad():
        for v in voxels:
            outs.append(model(v)['image'].clone().numpy())
    return np.stack(outs)


def _small_voxels(seed, frames, N, H, W):
    from utils.event_utils import events_to_voxel_torch
    vs = []
    for f in range(frames):
        batch = [events_to_voxel_torch(*[torch.from_numpy(a) for a in gen_events(seed * 100 + f * 10 + b, 600, H, W)],
                                       5, sensor_size=(H, W)) for b in range(N)]
        vs.append(torch.stack(batch))
    return vs


def _randomize_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(1.0 + 0.2 * torch.randn(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))


def golden_networks():
    import model as model_arch
    out = {}

    def save_model(tag, model, voxels, prefix):
        model.eval()
        frames = _run_frames(model, voxels)
        out[f'{tag}.voxels'] = torch.stack(voxels).numpy()
        out[f'{tag}.frames'] = frames
        for k, v in model.state_dict().items():
            if not k.endswith('num_batches_tracked'):
                out[f'{tag}.w.{k}'] = v.numpy()
        print(tag, 'frames', frames.shape, 'mean %.5f' % frames.mean())

    # real FireNet checkpoint (37k parameters) on a padded 48x64 input, batch 1, 4 frames
    ck = torch.load(os.path.join(REF, 'pretrained/FireNet/model.pth'), map_location='cpu')
    kw = dict(ck['config']['model'])
    kw['final_activation'] = ''
    m = model_arch.FireNet_legacy(kw)
    m.load_state_dict(ck['state_dict'])
    save_model('firenet_ckpt', m, _small_voxels(1, 4, 1, 48, 64), 'net.')

    # real FireNet+ checkpoint
    ck = torch.load(os.path.join(REF, 'pretrained/FireNet+/model.pth'), map_location='cpu')
    m = ck['config'].init_obj('arch', model_arch)
    m.load_state_dict(ck['state_dict'])
    save_model('firenetplus_ckpt', m, _small_voxels(2, 3, 2, 40, 56), '')

    # E2VID topology (BN, sigmoid) at base width 8, random weights through the real class, batch 2
    torch.manual_seed(10)
    kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
          'base_num_channels': 8, 'num_residual_blocks': 2, 'use_upsample_conv': True, 'norm': 'BN',
          'final_activation': 'sigmoid'}
    m = model_arch.E2VIDRecurrent(dict(kw))
    _randomize_bn(m, 11)
    save_model('e2vid_small', m, _small_voxels(3, 4, 2, 32, 48), 'unetrecurrent.')

    # E2VID+ topology (FlowNet, no norm, 3 output channels)
    torch.manual_seed(12)
    kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
          'base_num_channels': 4, 'num_residual_blocks': 2, 'use_upsample_conv': True, 'norm': 'none',
          'num_output_channels': 3}
    m = model_arch.FlowNet(dict(kw))
    save_model('flownet_small', m, _small_voxels(4, 3, 1, 32, 40), 'unetflow.')

    # HyperE2VID topology (dynamic decoder) at base width 4 (bottleneck 32 channels)
    torch.manual_seed(13)
    kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'kernel_size': 5,
          'channel_multiplier': 2, 'num_encoders': 3, 'base_num_channels': 4, 'num_residual_blocks': 2,
          'use_upsample_conv': True, 'norm': 'none', 'num_output_channels': 1, 'use_dynamic_decoder': True}
    m = model_arch.E2VIDRecurrent(dict(kw))
    _randomize_bn(m, 14)
    save_model('hyper_small', m, _small_voxels(5, 4, 2, 32, 48), 'unetrecurrent.')

    # TransposedConvLayer decoders (use_upsample_conv=False; model/submodules.py:38-66) -- no shipped checkpoint uses
    # them, so random weights through the real class: once with BN, once without
    torch.manual_seed(15)
    kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
          'base_num_channels': 8, 'num_residual_blocks': 2, 'use_upsample_conv': False, 'norm': 'BN',
          'final_activation': 'sigmoid'}
    m = model_arch.E2VIDRecurrent(dict(kw))
    _randomize_bn(m, 16)
    save_model('e2vid_tconv', m, _small_voxels(6, 3, 2, 32, 48), 'unetrecurrent.')
    np.savez_compressed(os.path.join(OUT, 'networks.npz'), **out)


def golden_networks_base32():
    """E2VID family at the SHIPPED width (base 32, last decoder 64 -> 32: the phase-stacked decoder of poly.cu) but with one
    encoder / one residual block so the fixture stays ~2 MB; sizes that are not multiples of the 8x16 GEMM tile."""
    import model as model_arch
    out = {}
    for tag, seed, norm, act, H, W in (('e2vid_b32', 20, 'BN', 'sigmoid', 30, 44), ('e2vid_b32_nonorm', 22, 'none', '', 18, 26)):
        torch.manual_seed(seed)
        kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 1,
              'base_num_channels': 32, 'num_residual_blocks': 1, 'use_upsample_conv': True, 'norm': norm,
              'final_activation': act}
        m = model_arch.E2VIDRecurrent(dict(kw))
        _randomize_bn(m, seed + 1)
        m.eval()
        voxels = _small_voxels(seed, 3, 2, H, W)
        frames = _run_frames(m, voxels)
        out[f'{tag}.voxels'] = torch.stack(voxels).numpy()
        out[f'{tag}.frames'] = frames
        for k, v in m.state_dict().items():
            if not k.endswith('num_batches_tracked'):
                out[f'{tag}.w.{k}'] = v.numpy()
        print(tag, 'frames', frames.shape, 'mean %.5f' % frames.mean())
    np.savez_compressed(os.path.join(OUT, 'networks_base32.npz'), **out)


def golden_real_slices():
    """REAL pretrained weights at the shipped width through the REAL classes (weak point of round 1: every full-width test
    used seeded weights).  A whole 43 MB checkpoint cannot live in git, but a sub-network can: an E2VIDRecurrent with TWO
    encoders and no residual block is a legitimate instance of the reference class whose every layer exists in
    pretrained/E2VID (head, encoders.0/1 with their ConvLSTM gates, decoders.1/2 renamed decoders.0/1, pred; eval-mode
    BatchNorm with the real running statistics), ~8 MB; the E2VID+ (FlowNet, no norm, 3 output channels) slice uses one
    encoder.  Plus one real 256-channel residual convolution on a real bottleneck activation captured with a forward hook
    in a full E2VID run (layer-level test through evk_conv2d_nhwc)."""
    import model as model_arch
    out = {}

    def slice_sd(sd, prefix, n_enc):
        keep = {}
        for k, v in sd.items():
            if k.endswith('num_batches_tracked'):
                continue
            name = k[len(prefix):]
            if name.startswith('head.') or name.startswith('pred.'):
                keep[k] = v
            elif name.startswith('encoders.'):
                if int(name.split('.')[1]) < n_enc:
                    keep[k] = v
            elif name.startswith('decoders.'):
                i = int(name.split('.')[1])
                j = i - (3 - n_enc)              # the LAST n_enc decoders of the real net are the decoders of the slice
                if j >= 0:
                    keep[prefix + 'decoders.%d.' % j + name.split('.', 2)[2]] = v
        return keep

    ck = torch.load(os.path.join(REF, 'pretrained/E2VID/model.pth'), map_location='cpu')
    kw = dict(ck['model'])
    kw.update(final_activation='sigmoid', num_encoders=2, num_residual_blocks=0)
    m = model_arch.E2VIDRecurrent(dict(kw))
    sd = slice_sd(ck['state_dict'], 'unetrecurrent.', 2)
    missing = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if 'num_batches_tracked' not in k], missing
    m.eval()
    voxels = _small_voxels(30, 3, 2, 44, 60)            # not a multiple of the 8x16 GEMM tile
    out['e2vid_real2.voxels'] = torch.stack(voxels).numpy()
    out['e2vid_real2.frames'] = _run_frames(m, voxels)
    for k, v in sd.items():
        out['e2vid_real2.w.' + k] = v.numpy()
    print('e2vid_real2 frames mean %.5f' % out['e2vid_real2.frames'].mean())

    ck = torch.load(os.path.join(REF, 'pretrained/E2VID+/model.pth'), map_location='cpu')
    kw = dict(ck['config']['arch']['args']['unet_kwargs'])
    kw.update(num_encoders=1, num_residual_blocks=0)
    m = model_arch.FlowNet(dict(kw))
    sd = slice_sd(ck['state_dict'], 'unetflow.', 1)
    missing = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if 'num_batches_tracked' not in k], missing
    m.eval()
    voxels = _small_voxels(31, 3, 2, 30, 44)
    out['flownet_real1.voxels'] = torch.stack(voxels).numpy()
    out['flownet_real1.frames'] = _run_frames(m, voxels)
    out['flownet_real1.kwargs'] = np.array([kw['base_num_channels'], kw['num_output_channels']], dtype=np.int64)
    for k, v in sd.items():
        out['flownet_real1.w.' + k] = v.numpy()
    print('flownet_real1 frames mean %.5f' % out['flownet_real1.frames'].mean())

    # layer level: resblocks.0 conv1 + bn1 + ReLU of the real E2VID on the real bottleneck activation
    ck = torch.load(os.path.join(REF, 'pretrained/E2VID/model.pth'), map_location='cpu')
    kw = dict(ck['model'])
    kw['final_activation'] = 'sigmoid'
    full = model_arch.E2VIDRecurrent(kw)
    full.load_state_dict(ck['state_dict'])
    full.eval()
    cap = {}
    rb = full.unetrecurrent.resblocks[0]
    h1 = rb.conv1.register_forward_hook(lambda mod, i, o: cap.__setitem__('x', i[0].detach().clone()))
    h2 = rb.bn1.register_forward_hook(lambda mod, i, o: cap.__setitem__('y', torch.relu(o.detach().clone())))
    with torch.no_grad():
        for v in _small_voxels(32, 2, 1, 64, 80):
            full(v)
    h1.remove(); h2.remove()
    out['resconv.x'] = cap['x'].numpy()                  # [1, 256, 8, 10]
    out['resconv.y'] = cap['y'].numpy()
    out['resconv.weight'] = rb.conv1.weight.detach().numpy()
    for k in ('weight', 'bias', 'running_mean', 'running_var'):
        out['resconv.bn.' + k] = getattr(rb.bn1, k).detach().numpy()
    np.savez_compressed(os.path.join(OUT, 'real_slices.npz'), **out)


def golden_full_checkpoints():
    """(Input grids are not stored: the test rebuilds them from gen_events(40 + f, 15000 + 7000 f) with the pinned oracle
    voxelizer and checks them against the stored sums.)
    The shipped 43 MB checkpoints cannot be committed, but they CAN travel to the GPU box as untracked files: this copies
    them to tests/golden/_ckpt/ (git-ignored, not gpurun-ignored) and commits what the REAL classes produce from them on
    CPU -- three recurrent frames at the BASELINE sizes -- so tests/test_gpu_checkpoints.py runs the real checkpoints
    through the CUDA path at full size whenever the files are present."""
    import shutil
    import model as model_arch
    from utils.event_utils import events_to_voxel_torch
    from utils.util import CropParameters
    import eval as ref_eval
    out = {}
    ckdir = os.path.join(OUT, '_ckpt')
    os.makedirs(ckdir, exist_ok=True)
    for name, (H, W, norm_ev) in {'E2VID': (180, 240, True), 'E2VID+': (180, 240, False), 'HyperE2VID': (260, 346, False),
                                  'SSL-E2VID': (180, 240, False), 'FireNet': (180, 240, True), 'FireNet+': (180, 240, False)}.items():
        src = os.path.join(REF, f'pretrained/{name}/model.pth')
        shutil.copyfile(src, os.path.join(ckdir, name + '.pth'))
        ck = torch.load(src, map_location='cpu')
        if name == 'E2VID':
            kw = dict(ck['model']); kw['final_activation'] = 'sigmoid'
            m = model_arch.E2VIDRecurrent(kw); sd = ck['state_dict']
        elif name == 'SSL-E2VID':
            kw = {"base_num_channels": 32, "kernel_size": 5, "num_bins": 5, "num_encoders": 3, "recurrent_block_type": "convlstm",
                  "num_residual_blocks": 2, "skip_type": "sum", "norm": None, "use_upsample_conv": True}
            m = model_arch.E2VIDRecurrent(kw); sd = ck
        elif name == 'FireNet':
            kw = dict(ck['config']['model']); kw['final_activation'] = ''
            m = model_arch.FireNet_legacy(kw); sd = ck['state_dict']
        else:
            m = ck['config'].init_obj('arch', model_arch); sd = ck['state_dict']
            if name == 'FireNet+':
                m.num_encoders = 0
        m.load_state_dict(sd)
        m.eval()
        cp = CropParameters(W, H, m.num_encoders)
        frames, voxels = [], []
        m.reset_states()
        with torch.no_grad():
            for f in range(3):
                v = events_to_voxel_torch(*[torch.from_numpy(a) for a in gen_events(40 + f, 15000 + 7000 * f, H, W)], 5, sensor_size=(H, W))[None]
                voxels.append(v[0].numpy())
                if norm_ev:
                    v = ref_eval.normalize_event_tensor(v)
                frames.append(cp.crop(m(cp.pad(v))['image'])[0, 0].numpy().copy())
        out[f'{name}.voxel_sums'] = np.array([[v.sum(dtype=np.float64), np.abs(v).sum(dtype=np.float64)] for v in voxels])   # (the test regenerates
        out[f'{name}.frames'] = np.stack(frames).astype(np.float32)
        out[f'{name}.meta'] = np.array([H, W, int(norm_ev), m.num_encoders], dtype=np.int64)
        print(name, 'frames mean', [float(f.mean()) for f in frames])
    np.savez_compressed(os.path.join(OUT, 'full_checkpoints.npz'), **out)


def golden_metrics():
    out = {}
    yy, xx = np.mgrid[0:180, 0:240]
    ref = (0.5 + 0.4 * np.sin(xx / 9) * np.cos(yy / 7)).astype(np.float32)
    img = np.clip(ref + np.random.default_rng(0).normal(0, 0.05, (180, 240)), 0, 1).astype(np.float32)
    out['a7.ref'], out['a7.img'] = ref, img
    # independent path: scipy.ndimage.gaussian_filter itself (what scikit-image calls)
    from scipy.ndimage import gaussian_filter

    def ssim_scipy(x, y):
        f = lambda a: gaussian_filter(a, 1.5, truncate=3.5, mode='reflect')
        ux, uy = f(x), f(y)
        vx, vy, vxy = f(x * x) - ux * ux, f(y * y) - uy * uy, f(x * y) - ux * uy
        C1, C2 = 0.01 ** 2, 0.03 ** 2
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        return float(S[5:-5, 5:-5].mean(dtype=np.float64))

    out['a7.scores'] = np.array([np.mean((ref - img) ** 2, dtype=np.float64), ssim_scipy(ref, img)])
    g = np.random.default_rng(7)
    pairs = []
    for i, (H, W) in enumerate([(180, 240), (130, 173), (37, 53), (11, 11), (64, 80)]):
        a = g.random((H, W)).astype(np.float32)
        b = np.clip(a + g.normal(0, 0.1 * (i + 1), (H, W)), 0, 1).astype(np.float32)
        out[f'pair{i}.img'], out[f'pair{i}.ref'] = a, b
        out[f'pair{i}.scores'] = np.array([np.mean((b - a) ** 2, dtype=np.float64), ssim_scipy(b, a)])
    # robust normalisation through the reference's own helper
    from utils.eval_utils import normalize
    x = (g.random((180, 240)) ** 2).astype(np.float32)
    out['norm.in'] = x
    out['norm.robust'] = normalize(x, 1, 99).astype(np.float32)
    out['norm.standard'] = normalize(x, 0, 100).astype(np.float32)
    out['norm.exprobust'] = normalize(np.exp(x), 1, 99).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'metrics.npz'), **out)
    print('metrics a7', out['a7.scores'])


def golden_eval_loop(tmp):
    """The REAL eval.eval_method_on_sequence on a small synthetic sequence with the real FireNet checkpoint
    (event_tensor_normalization on, no post-norm) and with an E2VID-topology model ('robust' post-norm)."""
    import eval as ref_eval
    import model as model_arch
    from dataset import MemMapDataset
    from torch.utils.data import DataLoader
    ref_eval.CudaTimer = lambda *_a, **_k: contextlib.nullcontext()
    ref_eval.tqdm = lambda x, *a, **k: x
    seq_path = synthetic.write_sequence(os.path.join(tmp, 'evalseq'), 48, 64, 60000.0, 1.0, 20.0, seed=9)
    out = {k: np.load(os.path.join(seq_path, k + '.npy')) for k in
           ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices')}
    eval_config = {'name': 'std', 'save_images': False, 'histeq': 'none', 'eval_infer_all': False, 'ts_tol_ms': 1.0,
                   'create_video': False}
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        def run(tag, model, method_config, start, end):
            ds = MemMapDataset(seq_path, num_bins=5, voxel_method={'method': 'between_frames'})
            sequence = {'name': 'evalseq', 'data_loader': DataLoader(ds), 'start_time_s': start, 'end_time_s': end}
            # capture per-frame scores by wrapping the tracker factory
            holder = {}
            orig = ref_eval.get_eval_metrics_tracker

            def wrapped(*a, **k):
                holder['t'] = orig(*a, **k)
                return holder['t']
            ref_eval.get_eval_metrics_tracker = wrapped
            n_eval, means = ref_eval.eval_method_on_sequence('SYN', eval_config, tag, model, method_config, sequence,
                                                             ['mse', 'ssim'])
            ref_eval.get_eval_metrics_tracker = orig
            t = holder['t']
            out[f'{tag}.indices'] = np.array(t.quan_eval_indices, dtype=np.int64)
            out[f'{tag}.mse'] = np.array(t.metrics[0].scores)
            out[f'{tag}.ssim'] = np.array(t.metrics[1].scores)
            out[f'{tag}.summary'] = np.array([n_eval, means['mse'], means['ssim'], start, end])
            print(tag, n_eval, means)

        ck = torch.load(os.path.join(REF, 'pretrained/FireNet/model.pth'), map_location='cpu')
        kw = dict(ck['config']['model'])
        kw['final_activation'] = ''
        m = model_arch.FireNet_legacy(kw)
        ref_eval.load_model(m, ck['state_dict'])
        run('firenet', m, {'event_tensor_normalization': True, 'post_process_norm': 'none'}, 0.2, 0.9)

        torch.manual_seed(10)      # same instance as networks.npz 'e2vid_small'
        kw = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
              'base_num_channels': 8, 'num_residual_blocks': 2, 'use_upsample_conv': True, 'norm': 'BN',
              'final_activation': 'sigmoid'}
        m = model_arch.E2VIDRecurrent(dict(kw))
        _randomize_bn(m, 11)
        m.eval()
        for p_ in m.parameters():
            p_.requires_grad = False
        run('e2vid_small', m, {'event_tensor_normalization': True, 'post_process_norm': 'robust'}, 0.0, 2.0)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, 'eval_loop.npz'), **out)


def golden_eval_loop_modes(tmp):
    """The REAL eval.eval_method_on_sequence with 'k_events' and 't_seconds' windows (real FireNet checkpoint): the voxel
    timestamp is then the window's last event, the reference frame is the CLOSEST one (dataset.py:150-166) and the
    ts_tol_ms gate of utils/eval_metrics.py:258-262 actually rejects frames; empty windows get patched timestamps
    (dataset.py:59-71: the stream below has a silent gap)."""
    import eval as ref_eval
    import model as model_arch
    from dataset import MemMapDataset
    from torch.utils.data import DataLoader
    ref_eval.CudaTimer = lambda *_a, **_k: contextlib.nullcontext()
    ref_eval.tqdm = lambda x, *a, **k: x
    s = synthetic.make_stream(48, 64, 60000.0, 1.0, 20.0, seed=11)
    # a silent gap of 0.12 s -> empty 't_seconds' windows
    keep = ~((s['events_ts'] > 0.40) & (s['events_ts'] < 0.52))
    s['events_ts'], s['events_xy'], s['events_p'] = s['events_ts'][keep], s['events_xy'][keep], s['events_p'][keep]
    s['image_event_indices'] = np.clip(np.searchsorted(s['events_ts'], s['images_ts'][:, 0], 'right') - 1, 0,
                                       len(s['events_ts']) - 1).reshape(-1, 1).astype(np.int64)
    seq_path = os.path.join(tmp, 'modeseq')
    os.makedirs(seq_path, exist_ok=True)
    for k in ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices'):
        np.save(os.path.join(seq_path, k + '.npy'), s[k])
    import json
    json.dump({'sensor_resolution': [48, 64]}, open(os.path.join(seq_path, 'metadata.json'), 'w'))
    out = {k: s[k] for k in ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices')}
    ck = torch.load(os.path.join(REF, 'pretrained/FireNet/model.pth'), map_location='cpu')
    kw = dict(ck['config']['model'])
    kw['final_activation'] = ''
    m = model_arch.FireNet_legacy(kw)
    ref_eval.load_model(m, ck['state_dict'])
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        modes = {'k_events': ({'method': 'k_events', 'k': 2500, 'sliding_window_w': 500}, 8.0),
                 't_seconds': ({'method': 't_seconds', 't': 0.04, 'sliding_window_t': 0.01}, 6.0)}
        for tag, (vm, tol) in modes.items():
            eval_config = {'name': 'std', 'save_images': False, 'histeq': 'none', 'eval_infer_all': False, 'ts_tol_ms': tol,
                           'create_video': False}
            ds = MemMapDataset(seq_path, num_bins=5, voxel_method=dict(vm))
            sequence = {'name': 'modeseq', 'data_loader': DataLoader(ds), 'start_time_s': 0.1, 'end_time_s': 0.85}
            holder = {}
            orig = ref_eval.get_eval_metrics_tracker

            def wrapped(*a, **k):
                holder['t'] = orig(*a, **k)
                return holder['t']
            ref_eval.get_eval_metrics_tracker = wrapped
            n_eval, means = ref_eval.eval_method_on_sequence('SYN', eval_config, tag, m,
                                                             {'event_tensor_normalization': True, 'post_process_norm': 'none'},
                                                             sequence, ['mse', 'ssim'])
            ref_eval.get_eval_metrics_tracker = orig
            t = holder['t']
            out[f'{tag}.indices'] = np.array(t.quan_eval_indices, dtype=np.int64)
            out[f'{tag}.mse'] = np.array(t.metrics[0].scores)
            out[f'{tag}.ssim'] = np.array(t.metrics[1].scores)
            out[f'{tag}.summary'] = np.array([n_eval, means['mse'], means['ssim'], 0.1, 0.85, tol])
            out[f'{tag}.vm'] = np.array([vm.get('k', 0), vm.get('sliding_window_w', 0), vm.get('t', 0), vm.get('sliding_window_t', 0)], dtype=np.float64)
            out[f'{tag}.len'] = np.array([len(ds)], dtype=np.int64)
            print(tag, 'items', len(ds), 'evaluated', n_eval, means)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, 'eval_loop_modes.npz'), **out)


def golden_spade(copy_ckpt=True):
    """SPADE-E2VID (model/spade_e2v.py Unet6) through the REAL class: (1) seeded weights from evreal_b200.synthetic (the test
    regenerates them from the seed: 43 MB of weights are not stored), batch 1, 40x56, three recurrent frames -- the first
    one takes the normalised-event branch of x_org, the others the previous reconstruction; (2) the shipped checkpoint at
    180x240 (padded to 184x240 by CropParameters(num_encoders = 3), eval.py:131-132), three frames, .pth copied next to
    the other untracked checkpoints."""
    import shutil
    from model.spade_e2v import Unet6
    from utils.event_utils import events_to_voxel_torch
    from utils.util import CropParameters
    out = {}
    m = Unet6()
    m.load_state_dict(synthetic.spade_state_dict(7))
    m.eval()
    voxels = _small_voxels(60, 3, 1, 40, 56)
    m.reset_states()
    with torch.no_grad():
        out['seeded.frames'] = np.stack([m(v.clone())['image'].numpy().copy() for v in voxels])
    out['seeded.voxels'] = torch.stack(voxels).numpy()
    src = os.path.join(REF, 'pretrained/SPADE-E2VID/model.pth')
    if copy_ckpt:
        os.makedirs(os.path.join(OUT, '_ckpt'), exist_ok=True)
        shutil.copyfile(src, os.path.join(OUT, '_ckpt', 'SPADE-E2VID.pth'))
    m = Unet6()
    m.load_state_dict(torch.load(src, map_location='cpu'))
    m.eval()
    H, W = 180, 240
    cp = CropParameters(W, H, 3)
    frames, sums = [], []
    m.reset_states()
    with torch.no_grad():
        for f in range(3):
            v = events_to_voxel_torch(*[torch.from_numpy(a) for a in gen_events(40 + f, 15000 + 7000 * f, H, W)], 5, sensor_size=(H, W))[None]
            sums.append([float(v.sum(dtype=torch.float64)), float(v.abs().sum(dtype=torch.float64))])
            frames.append(cp.crop(m(cp.pad(v))['image'])[0, 0].numpy().copy())
    out['ckpt.frames'] = np.stack(frames).astype(np.float32)
    out['ckpt.voxel_sums'] = np.array(sums)
    print('spade seeded mean %.5f  ckpt frames mean' % out['seeded.frames'].mean(), [float(f.mean()) for f in frames])
    np.savez_compressed(os.path.join(OUT, 'spade.npz'), **out)


def golden_etnet(copy_ckpt=True):
    """ET-Net (model/eitr) through the REAL class: seeded weights from evreal_b200.synthetic (regenerated by the test), batch 2,
    40x56 (35 tokens), three recurrent frames; and the shipped checkpoint at 180x240 (690 tokens), three frames."""
    import shutil
    import model as model_arch
    from utils.event_utils import events_to_voxel_torch
    from utils.util import CropParameters
    out = {}
    m = model_arch.EITR({'num_bins': 5, 'norm': None})
    m.load_state_dict(synthetic.etnet_state_dict(9))
    m.eval()
    voxels = _small_voxels(70, 3, 2, 40, 56)
    out['seeded.frames'] = _run_frames(m, voxels)
    out['seeded.voxels'] = torch.stack(voxels).numpy()
    src = os.path.join(REF, 'pretrained/ET-Net/model.pth')
    if copy_ckpt:
        os.makedirs(os.path.join(OUT, '_ckpt'), exist_ok=True)
        shutil.copyfile(src, os.path.join(OUT, '_ckpt', 'ET-Net.pth'))
    ck = torch.load(src, map_location='cpu')
    m = ck['config'].init_obj('arch', model_arch)
    m.load_state_dict(ck['state_dict'])
    m.eval()
    H, W = 180, 240
    cp = CropParameters(W, H, 3)
    frames, sums = [], []
    m.reset_states()
    with torch.no_grad():
        for f in range(3):
            v = events_to_voxel_torch(*[torch.from_numpy(a) for a in gen_events(40 + f, 15000 + 7000 * f, H, W)], 5, sensor_size=(H, W))[None]
            sums.append([float(v.sum(dtype=torch.float64)), float(v.abs().sum(dtype=torch.float64))])
            frames.append(cp.crop(m(cp.pad(v))['image'])[0, 0].numpy().copy())
    out['ckpt.frames'] = np.stack(frames).astype(np.float32)
    out['ckpt.voxel_sums'] = np.array(sums)
    print('etnet seeded mean %.5f  ckpt frames mean' % out['seeded.frames'].mean(), [float(f.mean()) for f in frames])
    np.savez_compressed(os.path.join(OUT, 'etnet.npz'), **out)


def golden_colornet():
    """The REAL ColorNet (model/model.py:46-105) around the real FireNet checkpoint on a Bayer-sized input: merged colour
    frames over three recurrent steps (five batch-1 forwards per frame with swapped states in the reference)."""
    import model as model_arch
    ck = torch.load(os.path.join(REF, 'pretrained/FireNet/model.pth'), map_location='cpu')
    kw = dict(ck['config']['model'])
    kw['final_activation'] = ''
    m = model_arch.FireNet_legacy(kw)
    m.load_state_dict(ck['state_dict'])
    m.eval()
    from model.model import ColorNet
    cn = ColorNet(m)
    voxels = _small_voxels(50, 3, 1, 40, 56)
    cn.reset_states()
    frames = []
    with torch.no_grad():
        for v in voxels:
            frames.append(cn(v)['image'].numpy().copy())
    np.savez_compressed(os.path.join(OUT, 'colornet.npz'), voxels=torch.stack(voxels).numpy(), frames=np.stack(frames))
    print('colornet frames', np.stack(frames).shape, 'mean %.4f' % np.stack(frames).mean())


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    install_shims()
    torch.set_num_threads(1)          # single-threaded index_put_ -> bit-reproducible voxel grids
    if '--only-base32' in sys.argv:   # added later; leaves the other fixtures untouched
        golden_networks_base32()
        sys.exit(0)
    if '--only-spade' in sys.argv:
        golden_spade()
        sys.exit(0)
    if '--only-etnet' in sys.argv:
        golden_etnet()
        sys.exit(0)
    if '--round2' in sys.argv:        # round-2 fixtures only (the round-1 ones stay byte-identical)
        if '--only-color' not in sys.argv:
            golden_real_slices()
            golden_full_checkpoints()
            with tempfile.TemporaryDirectory() as tmp:
                golden_eval_loop_modes(tmp)
        golden_colornet()
        golden_spade()
        golden_etnet()
        sys.exit(0)
    with tempfile.TemporaryDirectory() as tmp:
        golden_voxel()
        golden_glue()
        golden_windows(tmp)
        golden_networks()
        golden_networks_base32()
        golden_real_slices()
        golden_full_checkpoints()
        golden_metrics()
        golden_eval_loop(tmp)
        golden_eval_loop_modes(tmp)
        golden_colornet()
        golden_spade()
        golden_etnet()
    for f in sorted(os.listdir(OUT)):
        print('%8.1f kB  %s' % (os.path.getsize(os.path.join(OUT, f)) / 1e3, f))
