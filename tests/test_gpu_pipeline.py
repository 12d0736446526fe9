"""GPU: the lock-step SequenceBatch (bench.py's loop) gives every sequence the frames and scores the per-sequence
loop (evaluate.eval_method_on_sequence == eval.py:189-246) gives it, in both input modes (events resident in HBM /
events in pinned host memory), and against the golden per-frame scores of the real reference loop."""
import numpy as np
import pytest
import torch

from helpers import golden, weights_of

pytestmark = pytest.mark.gpu

KW = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
      'base_num_channels': 8, 'num_residual_blocks': 2, 'use_upsample_conv': True, 'norm': 'BN',
      'final_activation': 'sigmoid'}


def _datasets(g, n):
    from evreal_b200.dataset import MemMapDataset
    base = {k: g[k] for k in ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices')}
    out = []
    for i in range(n):
        a = dict(base)
        if i > 0:                                   # a different stream: mirrored coordinates, flipped polarity
            xy = a['events_xy'].copy()
            xy[:, 0] = 63 - xy[:, 0]
            a['events_xy'] = xy
            a['events_p'] = 1 - a['events_p']
        a['sensor_resolution'] = [48, 64]
        out.append(MemMapDataset(a, num_bins=5, voxel_method={'method': 'between_frames'}, resident=False))
    return out


@pytest.mark.parametrize('resident', [True, False])
def test_lockstep_batch_matches_per_sequence_loop(resident):
    from evreal_b200 import E2VIDRecurrent
    from evreal_b200.pipeline import SequenceBatch
    g = golden('eval_loop')
    full, _ = weights_of(golden('networks'), 'e2vid_small', 'unetrecurrent.')
    dss = _datasets(g, 3)
    model = E2VIDRecurrent(dict(KW)).load_state_dict(full).to('cuda')
    batch = SequenceBatch(model, dss, True, 'robust', resident=resident)
    batch.reset()
    scores = []
    for idx in range(len(batch)):
        s, img, n_ev = batch.step(idx)
        torch.cuda.synchronize()
        scores.append(s.clone().cpu().numpy())
        assert tuple(img.shape) == (3, 1, 48, 64)
        assert batch.launches > 0
        assert (batch.h2d_bytes > 0) == (not resident) or n_ev == 0
    batch.check_bounds()
    scores = np.stack(scores)                       # [frames, B, 2]
    # stream 0 is the golden sequence: indices evaluated by the real loop are all frames inside [start, end]
    idxs = [int(i) for i in g['e2vid_small.indices']]
    assert np.allclose(scores[idxs, 0, 0], g['e2vid_small.mse'], rtol=1e-4, atol=0)
    assert np.allclose(scores[idxs, 0, 1], g['e2vid_small.ssim'], rtol=1e-4, atol=1e-7)
    # every stream alone (batch of 1) == its row in the batch
    for b in range(3):
        single = SequenceBatch(E2VIDRecurrent(dict(KW)).load_state_dict(full).to('cuda'), [dss[b]], True, 'robust',
                               resident=True)
        single.reset()
        for idx in range(len(single)):
            s, _, _ = single.step(idx)
            # float atomics in the voxelizer commute but do not associate: run-to-run differences of a few ulp
            assert np.allclose(s.cpu().numpy()[0], scores[idx, b], rtol=2e-5, atol=1e-9), (b, idx)


def test_out_of_sensor_events_raise_once_per_sequence():
    from evreal_b200 import E2VIDRecurrent
    from evreal_b200.dataset import MemMapDataset
    from evreal_b200.pipeline import SequenceBatch
    g = golden('eval_loop')
    full, _ = weights_of(golden('networks'), 'e2vid_small', 'unetrecurrent.')
    a = {k: g[k] for k in ('events_ts', 'events_xy', 'events_p', 'images', 'images_ts', 'image_event_indices')}
    xy = a['events_xy'].copy()
    xy[5000, 0] = 64                                 # x == W
    a['events_xy'] = xy
    a['sensor_resolution'] = [48, 64]
    ds = MemMapDataset(a, num_bins=5, resident=False)
    batch = SequenceBatch(E2VIDRecurrent(dict(KW)).load_state_dict(full).to('cuda'), [ds], True, 'robust')
    batch.reset()
    for idx in range(4):
        batch.step(idx)
    with pytest.raises(IndexError):
        batch.check_bounds()
