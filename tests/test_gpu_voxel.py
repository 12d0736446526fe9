"""GPU parity: the CUDA voxelizer + glue kernels (through the C ABI) against the reference's golden vectors
(tests/golden/voxel.npz, glue.npz: produced by the real utils.event_utils / eval.normalize_event_tensor /
utils.util.CropParameters) and against the CPU oracle on seeded inputs.

Tolerance: the reference's own index_put_(accumulate=True) is not bit-reproducible for large windows
(SURVEY 8a); float atomics commute but do not associate, so the bar is |diff| <= 1e-5 * max(1, |grid|max)
for grids and exact equality wherever at most one event lands per cell."""
import numpy as np
import pytest
import torch

from helpers import golden, gen_events

pytestmark = pytest.mark.gpu

CASES = ['small', 'cfg1', 'bins3', 'one_event', 'two_equal_t', 'three_equal_t', 'negative_wrap', 'same_pixel']


def _tol(ref):
    return 1e-5 * max(1.0, float(np.abs(ref).max()))


@pytest.mark.parametrize('name', CASES)
def test_voxelizer_matches_reference_golden(name):
    from evreal_b200 import events_to_voxel_torch
    g = golden('voxel')
    H, W, bins = (int(v) for v in g[name + '.meta'])
    xs, ys, ts, ps = (torch.from_numpy(g['%s.%s' % (name, k)]) for k in ('xs', 'ys', 'ts', 'ps'))
    grid = events_to_voxel_torch(xs, ys, ts, ps, bins, sensor_size=(H, W))
    assert grid.is_cuda and grid.dtype == torch.float32 and tuple(grid.shape) == (bins, H, W)
    ref = g[name + '.grid']
    assert np.max(np.abs(grid.cpu().numpy() - ref)) <= _tol(ref)


def test_voxelizer_survey_known_answers():
    """seed 0, n=15000, 240x180 -> sum -6.0, sum|v| 14307.22."""
    from evreal_b200 import events_to_voxel_torch
    xs, ys, ts, ps = (torch.from_numpy(a) for a in gen_events(0, 15000, 180, 240))
    grid = events_to_voxel_torch(xs, ys, ts, ps, 5, sensor_size=(180, 240)).double()
    assert abs(float(grid.sum()) + 6.0) < 1e-2
    assert abs(float(grid.abs().sum()) - 14307.22) < 0.5


@pytest.mark.parametrize('n', [1, 2, 3, 5, 63, 64, 65, 255, 1000, 4099])
@pytest.mark.parametrize('offset', [0, 1, 3])
def test_voxelizer_ragged_and_misaligned_windows(n, offset):
    """Windows are arbitrary slices of the resident stream: every length / 16-byte phase must agree with the oracle."""
    from evreal_b200 import _lib
    from oracle import event_voxel as ov
    H, W, bins = 20, 28, 5
    xs, ys, ts, ps = gen_events(100 + n, n + offset, H, W)
    dev = [torch.from_numpy(a).cuda() for a in (xs, ys, ts, ps)]
    sl = [d[offset:] for d in dev]
    grid = torch.empty((bins, H, W), dtype=torch.float32, device='cuda')
    lib = _lib.load()
    _lib.check(lib.evk_voxelize(*[_lib.ptr(s) for s in sl], n, bins, H, W, _lib.ptr(grid), None, _lib.stream_ptr()))
    ref = ov.events_to_voxel_oracle(*[torch.from_numpy(a[offset:]) for a in (xs, ys, ts, ps)], bins, (H, W)).numpy()
    assert np.max(np.abs(grid.cpu().numpy() - ref)) <= _tol(ref)


def test_voxelizer_raw_format_matches_oracle():
    """evk_voxelize_raw fuses dataset.py:222-228/:52-58 (f64 subtract, round to f32, p*2-1)."""
    from evreal_b200 import events_to_voxel_raw
    from oracle import event_voxel as ov
    g = np.random.default_rng(5)
    n, H, W = 30000, 180, 240
    xy = np.stack([g.integers(0, W, n), g.integers(0, H, n)], 1).astype(np.int16)
    t = 1234.5 + np.sort(g.uniform(0, 0.04, n))
    p = g.integers(0, 2, n).astype(np.uint8)
    for i0, i1 in [(0, n), (17, 20011), (5, 6), (101, 103)]:
        grid = events_to_voxel_raw(xy[i0:i1], t[i0:i1], p[i0:i1], 5, sensor_size=(H, W)).cpu().numpy()
        xs, ys, ts, ps = ov.raw_window_to_f32(xy[i0:i1], t[i0:i1], p[i0:i1])
        ref = ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in (xs, ys, ts, ps)], 5, (H, W)).numpy()
        assert np.max(np.abs(grid - ref)) <= _tol(ref), (i0, i1)


def test_voxelizer_errors_like_the_reference():
    from evreal_b200 import events_to_voxel_torch
    one = torch.zeros(1)
    with pytest.raises(IndexError):                      # x == W
        events_to_voxel_torch(torch.tensor([8.0]), one, one, one + 1, 5, sensor_size=(8, 8))
    with pytest.raises(IndexError):                      # y < -H
        events_to_voxel_torch(one, torch.tensor([-9.0]), one, one + 1, 5, sensor_size=(8, 8))
    with pytest.raises(IndexError):                      # empty window: ts[-1]
        events_to_voxel_torch(one[:0], one[:0], one[:0], one[:0], 5, sensor_size=(8, 8))
    with pytest.raises(AssertionError):                  # length mismatch (utils/event_utils.py:45)
        events_to_voxel_torch(torch.zeros(2), one, one, one, 5, sensor_size=(8, 8))


def test_events_to_image():
    from evreal_b200 import events_to_image_torch
    xs = torch.tensor([1.0, 1.0, 2.0])
    ys = torch.tensor([0.0, 0.0, 3.0])
    ps = torch.tensor([1.0, 1.0, -1.0])
    img = events_to_image_torch(xs, ys, ps, sensor_size=(4, 4)).cpu()
    ref = torch.zeros(4, 4)
    ref[0, 1] = 2.0
    ref[3, 2] = -1.0
    assert torch.equal(img, ref)


@pytest.mark.parametrize('n', [40_000, 4_000_000])
def test_voxelizer_full_size_properties(n):
    """BASELINE cfg 5 sizes (640x480): size-independent properties.
    (1) mass conservation: every event spreads total weight 1 over its two bins -> sum(grid) == sum(p);
    (2) linearity: grid(p) + grid(-p) == 0; grid of the two halves of the stream, voxelized with the full
        window's time base, is not available through the API, so check polarity split instead:
        grid(p) == grid(max(p,0)) + grid(min(p,0))."""
    from evreal_b200 import events_to_voxel_torch
    H, W = 480, 640
    xs, ys, ts, ps = (torch.from_numpy(a).cuda() for a in gen_events(11, n, H, W, dur=0.04))
    grid = events_to_voxel_torch(xs, ys, ts, ps, 5, sensor_size=(H, W))
    assert abs(float(grid.double().sum()) - float(ps.double().sum())) < 1e-3 * max(1.0, n ** 0.5)
    neg = events_to_voxel_torch(xs, ys, ts, -ps, 5, sensor_size=(H, W))
    assert float((grid + neg).abs().max()) <= 1e-4
    pos = events_to_voxel_torch(xs, ys, ts, ps.clamp(min=0), 5, sensor_size=(H, W))
    ne = events_to_voxel_torch(xs, ys, ts, ps.clamp(max=0), 5, sensor_size=(H, W))
    assert float((grid - pos - ne).abs().max()) <= 1e-4 * max(1.0, float(pos.abs().max()))
    # every bin plane gets the triangular share of the mass: interior bins ~ n/4, edge bins ~ n/8 (uniform t)
    mass = pos.double().sum(dim=(1, 2)).cpu().numpy() / float(ps.clamp(min=0).sum())
    assert np.allclose(mass, [0.125, 0.25, 0.25, 0.25, 0.125], atol=0.01)


def test_voxelizer_large_window_vs_oracle():
    from evreal_b200 import events_to_voxel_torch
    from oracle import event_voxel as ov
    H, W, n = 260, 346, 111_000                      # MVSEC-shape window (cfg 4)
    ev = gen_events(21, n, H, W, dur=1 / 45)
    grid = events_to_voxel_torch(*[torch.from_numpy(a) for a in ev], 5, sensor_size=(H, W)).cpu().numpy()
    ref = ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in ev], 5, (H, W)).numpy()
    assert np.max(np.abs(grid - ref)) <= _tol(ref)


def test_voxelizer_cfg5_top_point_vs_oracle():
    """BASELINE cfg 5 top point (640x480, 4 M events in one 40 ms window) against the CPU oracle (round-1 verdict, weak #4:
    the largest oracle comparison was 111 k events), in both input formats (f32 SoA and raw int16/f64/u8)."""
    from evreal_b200 import _lib, events_to_voxel_torch
    from oracle import event_voxel as ov
    H, W, n = 480, 640, 4_000_000
    g = np.random.default_rng(31)
    xy = np.stack([g.integers(0, W, n), g.integers(0, H, n)], axis=1).astype(np.int16)
    t = 7.25 + np.sort(g.uniform(0, 0.04, n))                     # absolute float64 seconds, like events_ts.npy
    p = g.integers(0, 2, n).astype(np.uint8)
    xs, ys, ts, ps = ov.raw_window_to_f32(xy, t, p)
    ref = ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in (xs, ys, ts, ps)], 5, (H, W)).numpy()
    grid = events_to_voxel_torch(*[torch.from_numpy(a) for a in (xs, ys, ts, ps)], 5, sensor_size=(H, W)).cpu().numpy()
    assert np.max(np.abs(grid - ref)) <= _tol(ref)
    raw = torch.empty((5, H, W), dtype=torch.float32, device='cuda')
    oob = torch.zeros(1, dtype=torch.int32, device='cuda')
    dxy, dt, dp = torch.from_numpy(xy).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda()
    _lib.check(_lib.load().evk_voxelize_raw(_lib.ptr(dxy), _lib.ptr(dt), _lib.ptr(dp), n, 5, H, W, _lib.ptr(raw), _lib.ptr(oob),
                                            _lib.stream_ptr()))
    assert int(oob.item()) == 0
    assert np.max(np.abs(raw.cpu().numpy() - ref)) <= _tol(ref)


def test_normalize_pad_crop_match_reference_golden():
    from evreal_b200 import CropParameters, normalize_event_tensor
    from evreal_b200.util import normalize_pad
    g = golden('glue')
    for name in ('e2vid_180x240', 'firenet_180x240', 'mvsec_260x346', 'odd_37x53'):
        H, W, enc, Hc, Wc, top, left, iy0, iy1, ix0, ix1 = (int(v) for v in g[name + '.meta'])
        cp = CropParameters(W, H, enc)
        assert (cp.height_crop_size, cp.width_crop_size, cp.padding_top, cp.padding_left, cp.iy0, cp.iy1, cp.ix0,
                cp.ix1) == (Hc, Wc, top, left, iy0, iy1, ix0, ix1)
    v = torch.from_numpy(g['norm.in']).cuda()
    vn = normalize_event_tensor(v[None])
    assert np.max(np.abs(vn.cpu().numpy() - g['norm.out'])) <= 2e-6 * np.abs(g['norm.out']).max()
    cp = CropParameters(53, 37, 2)
    fused = normalize_pad(v[None], cp.height_crop_size, cp.width_crop_size, True)
    assert np.max(np.abs(fused.cpu().numpy() - g['norm.padded'])) <= 2e-6 * np.abs(g['norm.padded']).max()
    padded = cp.pad(torch.from_numpy(g['norm.out']).cuda())
    assert np.array_equal(padded.cpu().numpy(), g['norm.padded'])
    assert np.array_equal(cp.crop(padded).cpu().numpy(), g['norm.cropped_back'])


def test_normalize_all_zero_tensor_is_left_alone():
    from evreal_b200 import normalize_event_tensor
    z = torch.zeros(1, 5, 16, 16, device='cuda')
    assert float(normalize_event_tensor(z).abs().max()) == 0.0


@pytest.mark.parametrize('side', ['left', 'right'])
def test_device_searchsorted_matches_numpy(side):
    """evk_searchsorted_f64 (device-side boundaries for 't_seconds') against np.searchsorted, bit-exact:
    duplicates, queries equal to stored values, below the first / above the last timestamp, a one-element and an empty array."""
    from evreal_b200 import _lib
    lib = _lib.load()
    g = np.random.default_rng(5)
    t = np.sort(g.uniform(3.0, 4.0, 200_003))
    t[1000:1040] = t[1000]                                              # a run of equal timestamps
    q = np.concatenate([g.uniform(2.9, 4.1, 5000), t[::997], t[1000:1003], [t[0], t[-1], -1.0, 1e9,
                        np.nextafter(t[500], 0.0), np.nextafter(t[500], 10.0)]])
    for arr in (t, t[:1], t[:0]):
        td = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        qd = torch.from_numpy(q).cuda()
        out = torch.empty(len(q), dtype=torch.int64, device='cuda')
        _lib.check(lib.evk_searchsorted_f64(_lib.ptr(td) if len(arr) else None, len(arr), _lib.ptr(qd), len(q),
                                            1 if side == 'right' else 0, _lib.ptr(out), _lib.stream_ptr()))
        assert np.array_equal(out.cpu().numpy(), np.searchsorted(arr, q, side=side))


@pytest.mark.parametrize('mode', ['t_seconds', 't_seconds_sliding'])
def test_t_seconds_table_on_device_matches_real_dataset(tmp_path, mode):
    """The 't_seconds' window table computed from the RESIDENT timestamps equals the real MemMapDataset's
    (tests/golden/windows.npz, dataset.py:104-117) and the host table of the same object."""
    from evreal_b200.dataset import MemMapDataset
    from helpers import write_sequence_from_arrays
    vm = {'t_seconds': {'method': 't_seconds', 't': 0.04, 'sliding_window_t': 0.0},
          't_seconds_sliding': {'method': 't_seconds', 't': 0.05, 'sliding_window_t': 0.01}}[mode]
    g = golden('windows')
    arrays = {k: g[k] for k in ('events_ts', 'events_xy', 'events_p', 'images_ts', 'image_event_indices')}
    arrays['images'] = np.zeros((len(arrays['images_ts']), 32, 40, 1), dtype=np.uint8)
    path = write_sequence_from_arrays(str(tmp_path / 'seq'), arrays, (32, 40))
    ds = MemMapDataset(path, voxel_method=dict(vm), num_bins=5, resident=True)
    host_table = np.array(ds.event_indices, dtype=np.int64)
    assert np.array_equal(host_table, g[mode + '.table'])
    ds._upload()                                                        # timestamps resident: the table is rebuilt on the device
    assert np.array_equal(np.array(ds.compute_timeblock_indices(), dtype=np.int64), g[mode + '.table'])
