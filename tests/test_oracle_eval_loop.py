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


def test_lockstep_item_ranges_match_the_reference_loop():
    """evaluate.lockstep_item_range (the per-sequence item range and score gate of the lock-step form) against the item
    indices the REAL eval.eval_method_on_sequence evaluated (golden), plus the 10-second lead-in and eval_infer_all."""
    from evreal_b200.dataset import MemMapDataset
    from evreal_b200.evaluate import lockstep_item_range
    g = golden('eval_loop')
    ds = MemMapDataset(_arrays(g), num_bins=5, voxel_method={'method': 'between_frames'}, resident=False)
    ts = [float(ds.frame_ts[ds.window(i)[2]]) for i in range(len(ds))]
    for tag in ('firenet', 'e2vid_small'):
        _, _, _, start, end = g[tag + '.summary']
        first, count, gate = lockstep_item_range(ts, float(start), float(end))
        assert [first + k for k in range(count) if gate[k]] == list(g[tag + '.indices'])
        assert first == 0 and count == next((i for i, t in enumerate(ts) if t > end), len(ts))   # every item before the cut is reconstructed
    # lead-in: items more than 10 s before start_time_s are not reconstructed at all
    ts2 = [0.5 * i for i in range(60)]
    first, count, gate = lockstep_item_range(ts2, 20.0, 25.0)
    assert first == 20 and first + count - 1 == 50 and [first + k for k in range(count) if gate[k]] == list(range(40, 51))
    first, count, gate = lockstep_item_range(ts2, 20.0, 25.0, infer_all=True)
    assert (first, count) == (0, 60) and sum(gate) == 11
    assert lockstep_item_range([], 0.0, 1.0) == (0, 0, [])
