"""CPU: event-window index tables are bit-exact with the reference's MemMapDataset (host logic, no GPU)."""
import numpy as np
import pytest

from helpers import golden, write_sequence_from_arrays
from oracle import windows as ow

MODES = {
    'between_frames': {'method': 'between_frames'},
    'k_events': {'method': 'k_events', 'k': 1500, 'sliding_window_w': 0},
    'k_events_sliding': {'method': 'k_events', 'k': 1500, 'sliding_window_w': 500},
    't_seconds': {'method': 't_seconds', 't': 0.04, 'sliding_window_t': 0.0},
    't_seconds_sliding': {'method': 't_seconds', 't': 0.05, 'sliding_window_t': 0.01},
}


@pytest.fixture(scope='module')
def seq(tmp_path_factory):
    g = golden('windows')
    arrays = {k: g[k] for k in ('events_ts', 'events_xy', 'events_p', 'images_ts', 'image_event_indices')}
    arrays['images'] = np.zeros((len(arrays['images_ts']), 32, 40, 1), dtype=np.uint8)
    return write_sequence_from_arrays(str(tmp_path_factory.mktemp('seq')), arrays, (32, 40)), g


@pytest.mark.parametrize('mode', list(MODES))
def test_host_mirror_tables(seq, mode):
    from evreal_b200.dataset import MemMapDataset
    path, g = seq
    ds = MemMapDataset(path, voxel_method=dict(MODES[mode]), num_bins=5, resident=False)
    assert len(ds) == int(g[mode + '.len'][0])
    assert np.array_equal(np.array(ds.event_indices, dtype=np.int64), g[mode + '.table'])
    for row in g[mode + '.items']:
        i, idx0, idx1 = int(row[0]), int(row[1]), int(row[2])
        if idx0 < 0:
            with pytest.raises(ValueError):
                ds.window(i)
        else:
            w = ds.window(i)
            assert (int(w[0]), int(w[1])) == (idx0, idx1)
            assert max(int(w[1]) - int(w[0]), 0) == int(row[3])
    if mode != 'between_frames':
        ok = [r for r in g[mode + '.items'] if r[1] >= 0]
        assert [ds.get_closest_frame_index(r[4]) for r in ok] == list(g[mode + '.closest_frame'])


def test_oracle_tables(seq):
    _, g = seq
    t = g['events_ts']
    items = g['between_frames.items']
    assert ow.between_frames_windows(g['image_event_indices']) == [(int(r[1]), int(r[2])) for r in items]
    assert ow.between_frames_windows(g['image_event_indices'])[0][0] == ow.between_frames_windows(g['image_event_indices'])[0][1]
    assert np.array_equal(np.array(ow.k_events_windows(len(t), 1500, 0)), g['k_events.table'])
    assert np.array_equal(np.array(ow.k_events_windows(len(t), 1500, 500)), g['k_events_sliding.table'])
    assert np.array_equal(np.array(ow.t_seconds_windows(t, 0.04, 0.0)), g['t_seconds.table'])
    assert np.array_equal(np.array(ow.t_seconds_windows(t, 0.05, 0.01)), g['t_seconds_sliding.table'])


def test_empty_and_ragged_windows(tmp_path):
    from evreal_b200.dataset import MemMapDataset
    # 3 frames, the second one at the same event index as the first -> an empty window in the middle
    arrays = {'events_ts': np.linspace(0.0, 1.0, 100), 'events_xy': np.zeros((100, 2), dtype=np.int16),
              'events_p': np.zeros(100, dtype=np.uint8), 'images': np.zeros((4, 16, 16, 1), dtype=np.uint8),
              'images_ts': np.array([[0.2], [0.2], [0.7], [1.0]]),
              'image_event_indices': np.array([[19], [19], [69], [99]], dtype=np.int64)}
    path = write_sequence_from_arrays(str(tmp_path / 's'), arrays, (16, 16))
    ds = MemMapDataset(path, num_bins=5, resident=False)
    assert len(ds) == 3
    assert [tuple(int(v) for v in ds.window(i)[:2]) for i in range(3)] == [(19, 19), (19, 19), (19, 69)]
    assert ds.get_min_max_t() == (0.0, 1.0)


def test_shard_sequences_is_balanced_and_deterministic():
    from evreal_b200.evaluate import shard_sequences
    seqs = [{'name': 's%02d' % i} for i in range(16)]
    weights = [float((i * 37) % 11 + 1) for i in range(16)]
    for ws in (1, 2, 4, 8):
        shards = shard_sequences(seqs, weights, ws)
        assert sorted(sum(shards, [])) == list(range(16))
        loads = [sum(weights[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(weights)
        assert shards == shard_sequences(seqs, weights, ws)
