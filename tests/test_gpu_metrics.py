"""GPU parity: fused clip+MSE+SSIM kernel and the percentile normalisation kernel (through the C ABI) against
golden values (tests/golden/metrics.npz: scipy.ndimage path = what scikit-image calls; the reference's own
utils.eval_utils.normalize) and against the CPU oracle.

Tolerance: scores within 1e-4 relative (north_star); in practice the kernel follows scipy's summation order
(float64 accumulate per 1-D pass, float32 storage) so SSIM agrees to ~1e-7 and MSE to 1e-9 relative."""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu


def test_survey_a7_known_answer():
    from evreal_b200.eval_metrics import mse_ssim
    g = golden('metrics')
    s = mse_ssim(g['a7.img'], g['a7.ref']).cpu().numpy()[0]
    assert abs(s[0] - 0.00250339) < 1e-8 and abs(s[1] - 0.7030217) < 2e-6
    assert abs(s[0] - g['a7.scores'][0]) <= 1e-9 * g['a7.scores'][0]
    same = mse_ssim(g['a7.ref'], g['a7.ref']).cpu().numpy()[0]
    assert same[0] == 0.0 and abs(same[1] - 1.0) < 1e-7


@pytest.mark.parametrize('i', range(5))
def test_pairs_match_golden(i):
    from evreal_b200.eval_metrics import mse_ssim
    g = golden('metrics')
    img, ref, want = g['pair%d.img' % i], g['pair%d.ref' % i], g['pair%d.scores' % i]
    s = mse_ssim(img, ref).cpu().numpy()[0]
    assert abs(s[0] - want[0]) <= 1e-9 * want[0]
    assert abs(s[1] - want[1]) <= 1e-5 * abs(want[1]) + 1e-7, (s[1], want[1])


def test_batched_and_clip_against_oracle():
    from evreal_b200.eval_metrics import mse_ssim
    from oracle import metrics as om
    g = np.random.default_rng(12)
    imgs = (g.random((6, 260, 346)) * 1.4 - 0.2).astype(np.float32)       # outside [0,1]: exercises the fused clip
    refs = g.random((6, 260, 346)).astype(np.float32)
    got = mse_ssim(torch.from_numpy(imgs).cuda(), torch.from_numpy(refs).cuda(), clip=True).cpu().numpy()
    for k in range(6):
        a, b = np.clip(imgs[k], 0, 1), np.clip(refs[k], 0, 1)
        assert abs(got[k, 0] - om.mse_oracle(a, b)) <= 1e-9
        assert abs(got[k, 1] - om.ssim_oracle(a, b)) <= 1e-6


def test_metric_classes_keep_the_reference_contract():
    from evreal_b200.eval_metrics import MseMetric, SsimMetric
    g = golden('metrics')
    m, s = MseMetric(), SsimMetric()
    m.update(g['a7.img'], g['a7.ref'])
    s.update(g['a7.img'], g['a7.ref'])
    assert m.get_num_scores() == 1 and abs(m.get_last_score() - g['a7.scores'][0]) < 1e-10
    assert abs(s.get_mean_score() - g['a7.scores'][1]) < 2e-6
    assert MseMetric().get_mean_score() == -1
    with pytest.raises(ValueError):
        m.calculate(np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32))       # window 11 > image
    with pytest.raises(ValueError):
        m.calculate(np.zeros((20, 20), np.float32), np.zeros((20, 21), np.float32))


@pytest.mark.parametrize('norm', ['robust', 'standard', 'exprobust'])
def test_percentile_normalisation_matches_reference_helper(norm):
    from evreal_b200.eval_utils import post_process_normalization
    g = golden('metrics')
    x = torch.from_numpy(g['norm.in']).cuda()
    got = post_process_normalization(x, norm).cpu().numpy()
    ref = g['norm.' + norm]
    tol = 2e-6 if norm == 'exprobust' else 3e-7       # expf on the device vs np.exp: 1 ulp
    assert np.max(np.abs(got - ref)) <= tol * np.abs(ref).max()


def test_percentile_with_duplicates_and_odd_sizes():
    from evreal_b200.eval_utils import percentile_normalize
    from oracle import metrics as om
    g = np.random.default_rng(2)
    for shape in [(180, 240), (7, 13), (1, 101), (260, 346)]:
        x = np.round(g.normal(0, 1, shape), 1).astype(np.float32)          # heavy duplicates, negative values
        got = percentile_normalize(torch.from_numpy(x).cuda(), 1, 99).cpu().numpy()
        ref = om.robust_normalize_oracle(x, 1, 99)
        assert np.max(np.abs(got - ref)) <= 3e-7 * np.abs(ref).max(), shape
    xb = g.random((4, 50, 60)).astype(np.float32)
    got = percentile_normalize(torch.from_numpy(xb).cuda(), 1, 99, batched=True).cpu().numpy()
    for k in range(4):
        assert np.max(np.abs(got[k] - om.robust_normalize_oracle(xb[k], 1, 99))) <= 3e-7 * 1.1


def test_global_histogram_equalisation_vs_oracle():
    """hist_eq 'global' (utils/eval_metrics.py:326-331) on the GPU against the numpy restatement of skimage's algorithm."""
    from evreal_b200 import _lib
    from oracle import metrics as om
    g = np.random.default_rng(5)
    for H, W in ((180, 240), (37, 53)):
        img = np.clip(g.normal(0.45, 0.2, (H, W)), 0, 1).astype(np.float32)
        x = torch.from_numpy(img).cuda()
        out = torch.empty_like(x)
        _lib.check(_lib.load().evk_equalize_hist(_lib.ptr(x), _lib.ptr(out), 1, x.numel(), 0, _lib.stream_ptr()))
        ref = om.equalize_hist_oracle(img)
        assert np.max(np.abs(out.cpu().numpy() - ref)) <= 2e-6
    const = torch.full((16, 16), 0.25, device='cuda')
    out = torch.empty_like(const)
    _lib.check(_lib.load().evk_equalize_hist(_lib.ptr(const), _lib.ptr(out), 1, const.numel(), 0, _lib.stream_ptr()))
    assert np.allclose(out.cpu().numpy(), om.equalize_hist_oracle(const.cpu().numpy()))


def test_local_histogram_equalisation_vs_oracle():
    """hist_eq 'local' (utils/eval_metrics.py:332-339: rank.equalize over disk(55)) on the GPU against the numpy restatement of
    skimage's algorithm: bit-exact (integer counts, one double division); a small radius on a non-tile size, and the reference's
    radius 55 on an image smaller than the footprint (the population varies at every pixel)."""
    from evreal_b200 import _lib
    from oracle import metrics as om
    g = np.random.default_rng(9)
    for H, W, r in ((37, 53, 5), (48, 64, 55), (33, 20, 16)):
        img = np.clip(g.normal(0.45, 0.25, (H, W)), 0, 1).astype(np.float32)
        img[:3, :5] = 0.0
        img[-2:, :] = 1.0
        x = torch.from_numpy(img).cuda()
        out = torch.empty_like(x)
        _lib.check(_lib.load().evk_equalize_local(_lib.ptr(x), _lib.ptr(out), 1, H, W, r, 0, _lib.stream_ptr()))
        assert np.array_equal(out.cpu().numpy(), om.equalize_local_oracle(img, r)), (H, W, r)
    with pytest.raises(ValueError):
        _lib.check(_lib.load().evk_equalize_local(_lib.ptr(x), _lib.ptr(x), 1, H, W, r, 0, _lib.stream_ptr()))      # in place


def test_tracker_with_local_histeq_matches_oracle_scores():
    from evreal_b200.eval_metrics import EvalMetricsTracker
    from oracle import metrics as om
    g = golden('metrics')
    img, ref = g['a7.img'][:64, :80], g['a7.ref'][:64, :80]
    tr = EvalMetricsTracker(hist_eq='local', quan_eval_metric_names=['mse', 'ssim'], has_reference_frames=True, write_files=False)
    tr.update(0, torch.from_numpy(np.ascontiguousarray(img)).cuda(), torch.from_numpy(np.ascontiguousarray(ref)).cuda(), 0.5, 0.5)
    tr.finalize(0)
    a, b = om.equalize_local_oracle(np.clip(img, 0, 1)), om.equalize_local_oracle(np.clip(ref, 0, 1))
    means = tr.get_mean_scores()
    assert abs(means['mse'] - om.mse_oracle(a, b)) <= 1e-4 * om.mse_oracle(a, b)
    assert abs(means['ssim'] - om.ssim_oracle(a, b)) <= 1e-4


def test_tracker_with_global_histeq_matches_oracle_scores():
    from evreal_b200.eval_metrics import EvalMetricsTracker
    from oracle import metrics as om
    g = golden('metrics')
    img, ref = g['a7.img'], g['a7.ref']
    tr = EvalMetricsTracker(hist_eq='global', quan_eval_metric_names=['mse', 'ssim'], has_reference_frames=True, write_files=False)
    tr.update(0, torch.from_numpy(img).cuda(), torch.from_numpy(ref).cuda(), 0.5, 0.5)
    tr.finalize(0)
    a, b = om.equalize_hist_oracle(np.clip(img, 0, 1)), om.equalize_hist_oracle(np.clip(ref, 0, 1))
    means = tr.get_mean_scores()
    assert abs(means['mse'] - om.mse_oracle(a, b)) <= 1e-4 * om.mse_oracle(a, b)
    assert abs(means['ssim'] - om.ssim_oracle(a, b)) <= 1e-4


def test_uint8_quantise_kernel_is_round_half_even_after_clip():
    from evreal_b200 import _lib
    from oracle import metrics as om
    vals = np.concatenate([np.linspace(-0.2, 1.2, 3001), (np.arange(256) + 0.5) / 255.0, [0.5 / 255, 2.5 / 255]]).astype(np.float32)
    x = torch.from_numpy(vals).cuda()
    q = torch.empty(x.shape, dtype=torch.uint8, device='cuda')
    _lib.check(_lib.load().evk_quantize_u8(_lib.ptr(x), _lib.ptr(q), x.numel(), _lib.stream_ptr()))
    assert np.array_equal(q.cpu().numpy(), om.quantize_u8_oracle(vals))
