"""Oracle: event-window index tables (integer, bit-exact requirement).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/dataset.py:
  :104-117  compute_timeblock_indices  ('t_seconds': chained windows, np.searchsorted left)
  :119-130  compute_k_indices          ('k_events')
  :168-186  set_voxel_method           (dataset length per mode)
  :287-294  compute_frame_indices      ('between_frames' from image_event_indices)
  :33-46    __getitem__ window lookup  (between_frames item i = [end(i-1), end(i)), item 0 empty)
"""
import numpy as np


def frame_index_table(image_event_indices):
    """dataset.py:287-294 -> list of [start, end] per frame, chained."""
    table = []
    start = 0
    for row in np.asarray(image_event_indices).reshape(len(image_event_indices), -1):
        end = row[0]
        table.append([start, end])
        start = end
    return table


def between_frames_windows(image_event_indices):
    """Windows the dataset yields for items 0..num_frames-2 (dataset.py:35-43,182).

    item i uses prev = table[i-1] (or table[0] when i == 0) and cur = table[i];
    the window is [prev.end, cur.end) -- so item 0 is always empty and the
    last frame is never used."""
    table = frame_index_table(image_event_indices)
    out = []
    for i in range(len(table) - 1):
        prev = table[i - 1] if i > 0 else table[0]
        out.append((int(prev[1]), int(table[i][1])))
    return out


def k_events_windows(num_events, k, sliding_window_w):
    """dataset.py:119-130,174."""
    length = max(int(num_events / (k - sliding_window_w)), 0)
    return [((k - sliding_window_w) * i, (k - sliding_window_w) * i + k) for i in range(length)]


def t_seconds_windows(event_ts, t, sliding_window_t):
    """dataset.py:104-117,177-178.  f64 expression order kept:
    start_time = ((t - sw) * i) + t0 ; end_time = start_time + t."""
    event_ts = np.asarray(event_ts)
    t0, tk = event_ts[0], event_ts[-1]
    length = max(int((tk - t0) / (t - sliding_window_t)), 0)
    out = []
    start_idx = 0
    for i in range(length):
        start_time = ((t - sliding_window_t) * i) + t0
        end_time = start_time + t
        end_idx = int(np.searchsorted(event_ts, end_time))
        out.append((start_idx, end_idx))
        start_idx = end_idx
    return out
