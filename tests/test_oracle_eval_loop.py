"""CPU: the oracle's sequence loop against per-frame scores of the REAL eval.eval_method_on_sequence."""
import numpy as np

from helpers import golden, weights_of
from oracle import eval_loop, networks as on


def _arrays(g):
    return {k: g[k] for k in ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices')}


def test_firenet_real_checkpoint_sequence():
    g = golden('eval_loop')
    _, w = weights_of(golden('networks'), 'firenet_ckpt', 'net.')
    n_eval, mse, ssim, start, end = g['firenet.summary']
    res = eval_loop.run_sequence(_arrays(g), (48, 64), on.FireNetLegacyOracle(w), 4, True, 'none', start, end)
    assert res['indices'] == list(g['firenet.indices'])
    assert len(res['mse']) == int(n_eval)
    assert np.allclose(res['mse'], g['firenet.mse'], rtol=1e-6, atol=0)
    assert np.allclose(res['ssim'], g['firenet.ssim'], rtol=1e-5, atol=1e-8)   # conv summation order varies with the thread count
    assert abs(np.mean(res['mse']) - mse) < 1e-8 and abs(np.mean(res['ssim']) - ssim) < 1e-8


def test_e2vid_topology_sequence_with_robust_norm():
    g = golden('eval_loop')
    _, w = weights_of(golden('networks'), 'e2vid_small', 'unetrecurrent.')
    n_eval, mse, ssim, start, end = g['e2vid_small.summary']
    res = eval_loop.run_sequence(_arrays(g), (48, 64), on.UNetRecurrentOracle(w, final_sigmoid=True), 3, True, 'robust',
                                 start, end)
    assert res['indices'] == list(g['e2vid_small.indices'])
    assert np.allclose(res['mse'], g['e2vid_small.mse'], rtol=1e-5, atol=0)
    assert np.allclose(res['ssim'], g['e2vid_small.ssim'], rtol=1e-5, atol=1e-8)


def test_weighted_means():
    assert eval_loop.weighted_means([(2, {'mse': 1.0}), (0, {'mse': -1}), (6, {'mse': 3.0})]) == {'mse': 2.5}
