"""CPU: the oracle's voxelizer / glue against golden vectors produced by the real reference."""
import numpy as np
import torch

from helpers import golden
from oracle import event_voxel as ov

CASES = ['small', 'cfg1', 'bins3', 'one_event', 'two_equal_t', 'three_equal_t', 'negative_wrap', 'same_pixel']


def _case(g, name):
    H, W, bins = [int(v) for v in g[name + '.meta']]
    return [g[f'{name}.{k}'] for k in ('xs', 'ys', 'ts', 'ps')], (H, W), bins, g[name + '.grid']


def test_numpy_restatement_is_bit_exact():
    g = golden('voxel')
    for name in CASES:
        ev, size, bins, ref = _case(g, name)
        got = ov.events_to_voxel_numpy(*ev, bins, size)
        assert np.array_equal(got, ref), name


def test_torch_restatement_matches():
    g = golden('voxel')
    torch.set_num_threads(1)
    for name in CASES:
        ev, size, bins, ref = _case(g, name)
        got = ov.events_to_voxel_oracle(*[torch.from_numpy(v) for v in ev], bins, size).numpy()
        assert np.array_equal(got, ref), name


def test_survey_known_answers():
    g = golden('voxel')
    grid = g['cfg1.grid']
    assert abs(float(grid.sum()) - (-6.0)) < 1e-3
    assert abs(float(np.abs(grid).sum()) - 14307.22) < 0.05


def test_linspace_restatement():
    for steps in (1, 2, 3, 7, 64, 1000):
        ref = torch.linspace(0, 4, steps=steps).numpy()
        got = ov.linspace_f32(0, 4, steps)
        # ATen's vectorised kernel differs from the scalar formula by at most 1 ulp
        assert np.max(np.abs(ref - got)) <= 2.4e-7 * 4, steps


def test_out_of_range_raises_index_error():
    import pytest
    with pytest.raises(IndexError):
        ov.events_to_voxel_numpy([8], [0], [0.0], [1.0], 5, (8, 8))


def test_crop_parameters_and_normalisation():
    g = golden('glue')
    for name in ('e2vid_180x240', 'firenet_180x240', 'mvsec_260x346', 'odd_37x53'):
        H, W, enc, hp, wp, top, left, iy0, iy1, ix0, ix1 = [int(v) for v in g[name + '.meta']]
        c = ov.CropOracle(W, H, enc)
        assert (c.hp, c.wp, c.top, c.left, c.iy0, c.iy1, c.ix0, c.ix1) == (hp, wp, top, left, iy0, iy1, ix0, ix1)
    v = torch.from_numpy(g['norm.in'])
    vn = ov.normalize_event_tensor_oracle(v[None])
    assert np.array_equal(vn.numpy(), g['norm.out'])
    c = ov.CropOracle(53, 37, 2)
    assert np.array_equal(c.pad(vn).numpy(), g['norm.padded'])
    assert np.array_equal(c.crop(c.pad(vn)).numpy(), g['norm.cropped_back'])


def test_raw_casts():
    xy = np.array([[3, 4], [5, 6]], dtype=np.int16)
    t = np.array([10.000001, 10.5], dtype=np.float64)
    p = np.array([0, 1], dtype=np.uint8)
    xs, ys, ts, ps = ov.raw_window_to_f32(xy, t, p)
    assert xs.dtype == ys.dtype == ts.dtype == ps.dtype == np.float32
    assert ts[0] == 0 and ts[1] == np.float32(10.5 - 10.000001)
    assert list(ps) == [-1.0, 1.0]
