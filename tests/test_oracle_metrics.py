"""CPU: metric / normalisation oracle against golden values (scipy.ndimage path + reference's normalize)."""
import numpy as np

from helpers import golden
from oracle import metrics as om


def test_survey_a7_known_answer():
    g = golden('metrics')
    mse, ssim = om.mse_oracle(g['a7.img'], g['a7.ref']), om.ssim_oracle(g['a7.img'], g['a7.ref'])
    assert abs(ssim - 0.7030217) < 2e-7 and abs(mse - 0.00250339) < 1e-8
    assert abs(mse - g['a7.scores'][0]) < 1e-15 and abs(ssim - g['a7.scores'][1]) < 1e-12
    assert om.ssim_oracle(g['a7.ref'], g['a7.ref']) == 1.0


def test_pairs_match_scipy_gaussian_filter_path():
    g = golden('metrics')
    for i in range(5):
        img, ref, want = g[f'pair{i}.img'], g[f'pair{i}.ref'], g[f'pair{i}.scores']
        assert abs(om.mse_oracle(img, ref) - want[0]) < 1e-15
        assert abs(om.ssim_oracle(img, ref) - want[1]) < 1e-12, i


def test_gaussian_blur_matches_scipy_bitwise():
    from scipy.ndimage import gaussian_filter
    a = np.random.default_rng(3).random((50, 70)).astype(np.float32)
    assert np.array_equal(om.gaussian_blur_f32(a), gaussian_filter(a, 1.5, truncate=3.5, mode='reflect'))
    taps = om.gaussian_taps()
    assert len(taps) == 11 and abs(taps[5] - 0.26601172) < 1e-8 and abs(taps[0] - 0.00102838) < 1e-8


def test_ssim_symmetry_and_border_independence():
    g = np.random.default_rng(4)
    a, b = g.random((40, 44)).astype(np.float32), g.random((40, 44)).astype(np.float32)
    assert abs(om.ssim_oracle(a, b) - om.ssim_oracle(b, a)) < 1e-7


def test_percentile_normalisation_matches_reference_helper():
    g = golden('metrics')
    x = g['norm.in']
    assert np.array_equal(om.post_process_oracle(x, 'robust'), g['norm.robust'])
    assert np.array_equal(om.post_process_oracle(x, 'standard'), g['norm.standard'])
    assert np.allclose(om.post_process_oracle(x, 'exprobust'), g['norm.exprobust'], rtol=0, atol=1e-6)
    assert om.post_process_oracle(x, 'none') is x


def test_lpips_restatement_properties():
    import torch
    for net in ('alex', 'vgg'):
        w = om.random_lpips_weights(net, seed=1)
        a = torch.rand(2, 3, 64, 80, generator=torch.Generator().manual_seed(0))
        b = torch.rand(2, 3, 64, 80, generator=torch.Generator().manual_seed(1))
        assert float(om.lpips_oracle(a, a, w, net).abs().max()) == 0.0
        d = om.lpips_oracle(a, b, w, net)
        assert d.shape == (2,) and bool((d > 0).all())


def test_local_equalisation_oracle_known_answers():
    """rank.equalize restatement: a constant image maps to 255/255 everywhere (every neighbour <= g), a strictly increasing ramp
    with a footprint covering the whole image maps pixel k of n to floor(255 (k + 1) / n) / 255."""
    from oracle import metrics as om
    const = np.full((9, 11), 0.3, dtype=np.float32)
    assert np.array_equal(om.equalize_local_oracle(const, 3), np.ones((9, 11), dtype=np.float32))
    n = 12
    ramp = (np.arange(n, dtype=np.float32) * 20 / 255).reshape(1, n)
    got = om.equalize_local_oracle(ramp, 55)
    want = (np.floor(255 * (np.arange(n) + 1) / n).astype(np.uint8)).astype(np.float32) * np.float32(1 / 255.0)
    assert np.array_equal(got[0], want)
