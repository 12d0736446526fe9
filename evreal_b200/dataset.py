"""Host-side mirror of EVREAL's ``dataset.MemMapDataset`` (dataset.py:14-294).

Window index tables (``between_frames`` / ``k_events`` / ``t_seconds``), item
dictionary keys, timestamps and error behaviour follow the reference exactly
(integer index math is bit-exact, tests/test_windows.py).  What differs is where
the per-event work happens: with ``resident=True`` (default when CUDA is
available) the raw event arrays are uploaded to HBM once per sequence
(13 B/event: int16 x,y + float64 t + uint8 p) and every window is voxelized on
the GPU straight from those arrays (``evk_voxelize_raw`` fuses the casts of
dataset.py:222-228 and :52-58); the item's 'events' and 'frame' are CUDA tensors.
"""
import os
from bisect import bisect_left

import numpy as np
import torch

from . import _lib
from .event_utils import events_to_voxel_torch
from .util import read_json


class MemMapDataset(torch.utils.data.Dataset):

    def __init__(self, data_path, sensor_resolution=None, num_bins=5,
                 voxel_method=None, max_length=None, keep_ratio=1, device=None, resident=True):
        self.num_bins = num_bins
        self.data_path = data_path
        self.keep_ratio = keep_ratio
        self.sensor_resolution = sensor_resolution
        self.has_images = True
        self.channels = self.num_bins
        self.device = device
        self.resident = resident
        self._dev_events = None
        self._dev_images = None
        self.oob_total = None
        self.load_data(data_path)
        if voxel_method is None:
            voxel_method = {'method': 'between_frames'}
        self.voxel_method = voxel_method
        self.set_voxel_method()
        if max_length is not None:
            self.length = min(self.length, max_length + 1)

    # ------------------------------------------------------------------ windows
    def window(self, index):
        """[idx0, idx1) event range of item ``index`` (dataset.py:35-46)."""
        assert 0 <= index < self.__len__(), f"index {index} out of bounds (0 <= x < {self.__len__()})"
        if self.voxel_method['method'] == 'between_frames':
            prev_index = self.frames_to_use[index - 1] if index > 0 else 0
            frame_index = self.frames_to_use[index]
            _, idx0 = self.get_event_indices(prev_index)
            _, idx1 = self.get_event_indices(frame_index)
            return idx0, idx1, frame_index
        idx0, idx1 = self.get_event_indices(index)
        return idx0, idx1, index

    def item_meta(self, index):
        """Everything ``__getitem__`` returns except the tensors (dataset.py:33-102): (idx0, idx1, frame_index, voxel_timestamp,
        frame_timestamp, dt, event_count).  Host integer / float64 arithmetic only: the lock-step pipeline gates and pairs
        frames from these without voxelizing anything."""
        idx0, idx1, index = self.window(index)
        idx0, idx1 = int(idx0), int(idx1)
        event_count = max(idx1 - idx0, 0)
        t = self.filehandle["t"]
        method = self.voxel_method['method']
        if event_count > 0:
            ts_0, ts_k = t[idx0], t[idx1 - 1]
        elif idx0 > 0:
            # empty window: timestamps patched like dataset.py:59-71
            ts_0 = t[idx0 - 1:idx1][-1]
            ts_k = ts_0 + self.voxel_method['t'] if method == 't_seconds' else self.frame_ts[index]
        else:
            ts_0, ts_k = 0, 0
        dt = self.voxel_method['t'] if method == 't_seconds' else ts_k - ts_0
        if self.has_images and method != 'between_frames':
            index = self.get_closest_frame_index(ts_k)
        frame_timestamp = float(self.frame_ts[index]) if self.has_images else 0.0
        voxel_timestamp = frame_timestamp if method == 'between_frames' else float(ts_k)
        return idx0, idx1, index, voxel_timestamp, frame_timestamp, float(dt), event_count

    def __getitem__(self, index):
        idx0, idx1, index, voxel_timestamp, frame_timestamp, dt, event_count = self.item_meta(index)
        voxel = self.get_voxel_grid_window(idx0, idx1) if event_count > 0 else self.get_empty_voxel_grid()
        if self.has_images:
            frame = self.get_frame_tensor(index)
        else:
            frame = torch.zeros((1, self.sensor_resolution[0], self.sensor_resolution[1]),
                                dtype=torch.float32, device=voxel.device)
        return {'frame': frame,
                'events': voxel,
                'frame_timestamp': torch.tensor(frame_timestamp, dtype=torch.float64),
                'voxel_timestamp': torch.tensor(voxel_timestamp, dtype=torch.float64),
                'dt': torch.tensor(dt, dtype=torch.float64),
                'event_count': event_count}

    def compute_timeblock_indices(self):
        """'t_seconds' windows (dataset.py:104-117): window i ends at the first event at or after ((t - sw) * i + t0) + t --
        the reference's float64 expression order, evaluated for all i at once -- and starts where window i-1 ended."""
        t, sw = self.voxel_method['t'], self.voxel_method['sliding_window_t']
        end_times = ((t - sw) * np.arange(len(self), dtype=np.float64) + self.t0) + t
        if self._dev_events is not None and len(end_times) > 0:
            # the timestamps are already resident: search them where they are (evk_searchsorted_f64, float64 compares only)
            return self._chained(self.searchsorted_device(end_times))
        return self._chained(np.searchsorted(self.filehandle["t"], end_times, side='left'))

    def searchsorted_device(self, values, side='left'):
        """np.searchsorted(self.filehandle['t'], values, side) on the device-resident float64 timestamps."""
        self._upload()
        dev = self._device()
        t = self._dev_events[1]
        vals = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(dev)
        out = torch.empty(vals.numel(), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().evk_searchsorted_f64(_lib.ptr(t), t.numel(), _lib.ptr(vals), vals.numel(),
                                                        1 if side == 'right' else 0, _lib.ptr(out), _lib.stream_ptr(dev)))
        return out.cpu().numpy()

    def compute_k_indices(self):
        """'k_events' windows (dataset.py:119-130)."""
        k, w = self.voxel_method['k'], self.voxel_method['sliding_window_w']
        return [[(k - w) * i, (k - w) * i + k] for i in range(len(self))]

    def compute_frame_indices(self):
        """'between_frames' table (dataset.py:287-294): frame j closes the window that frame j-1 opened."""
        return self._chained(np.asarray(self.filehandle["image_event_indices"])[:, 0])

    @staticmethod
    def _chained(ends):
        """[[0, e0], [e0, e1], ...] for window end indices e0, e1, ..."""
        ends = [int(e) for e in ends]
        return [[a, b] for a, b in zip([0] + ends[:-1], ends)]

    def choose_frames_to_use(self):
        """keep_ratio < 1 evaluates a random subset of the frames (dataset.py:168-176; the draw is numpy's global
        generator, as in the reference, so a seeded run picks the same frames)."""
        every = list(range(self.num_frames))
        self.frames_to_use = every
        if self.keep_ratio == 1:
            return
        assert self.voxel_method['method'] == 'between_frames', "keep_ratio can only specified for between_frames voxel method"
        assert self.keep_ratio < 1, "keep_ratio cannot be greater than 1"
        kept = int(self.num_frames * self.keep_ratio)
        self.frames_to_use = sorted(np.random.choice(every, size=kept, replace=False))
        self.length = kept - 1

    def get_min_max_t(self):
        if self.has_images:
            return min(self.frame_ts[0], self.t0), max(self.frame_ts[-1], self.tk)
        return self.t0, self.tk

    def get_closest_frame_index(self, ts):
        pos = bisect_left(self.frame_ts, ts)
        if pos == 0:
            return 0
        if pos == len(self.frame_ts):
            return pos - 1
        before, after = self.frame_ts[pos - 1], self.frame_ts[pos]
        return pos if after - ts < ts - before else pos - 1

    def set_voxel_method(self):
        """Dataset length and event-index table of the configured windowing (dataset.py:178-186)."""
        vm = self.voxel_method
        kind = vm['method']
        if kind == 'between_frames':
            assert self.has_images, "Cannot use between_frames voxel method without images"
            self.length = self.num_frames - 1
            self.event_indices = self.compute_frame_indices()
            self.choose_frames_to_use()
            return
        if kind == 'k_events':
            span, total, table = vm['k'] - vm['sliding_window_w'], self.num_events, self.compute_k_indices
        elif kind == 't_seconds':
            span, total, table = vm['t'] - vm['sliding_window_t'], self.tk - self.t0, self.compute_timeblock_indices
        else:
            raise ValueError("Invalid voxel forming method chosen ({})".format(vm))
        self.length = max(int(total / span), 0)
        self.event_indices = table()

    def __len__(self):
        return self.length

    def get_event_indices(self, index):
        idx0, idx1 = self.event_indices[index]
        if not (idx0 >= 0 and idx1 <= self.num_events):
            raise ValueError("WARNING: Event indices {},{} out of bounds 0,{}".format(idx0, idx1, self.num_events))
        return idx0, idx1

    # ------------------------------------------------------------------ device side
    def _device(self):
        _lib.require_cuda()
        if self.device is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        return torch.device(self.device)

    def _upload(self):
        """One H2D copy of the raw arrays per sequence (pinned staging, 13 B/event)."""
        if self._dev_events is not None:
            return
        dev = self._device()
        fh = self.filehandle
        xy = np.ascontiguousarray(fh["xy"], dtype=np.int16)
        t = np.ascontiguousarray(fh["t"], dtype=np.float64)
        p = np.ascontiguousarray(fh["p"]).astype(np.uint8)
        self._dev_events = tuple(torch.from_numpy(a).pin_memory().to(dev, non_blocking=True) for a in (xy, t, p))
        self.oob_total = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.has_images:
            imgs = np.ascontiguousarray(fh["images"][..., 0])
            self._dev_images = torch.from_numpy(imgs).pin_memory().to(dev, non_blocking=True)

    def get_empty_voxel_grid(self):
        size = (self.num_bins, *self.sensor_resolution)
        return torch.zeros(size, dtype=torch.float32, device=self._device())

    def get_voxel_grid_window(self, idx0, idx1):
        """Voxel grid of events [idx0, idx1)."""
        H, W = int(self.sensor_resolution[0]), int(self.sensor_resolution[1])
        if self.resident:
            self._upload()
            dev = self._device()
            xy, t, p = self._dev_events
            grid = torch.empty((self.num_bins, H, W), dtype=torch.float32, device=dev)
            n = idx1 - idx0
            with torch.cuda.device(dev):
                _lib.check(_lib.load().evk_voxelize_raw(
                    _lib.ptr(xy[idx0:idx1]), _lib.ptr(t[idx0:idx1]), _lib.ptr(p[idx0:idx1]), n, self.num_bins, H, W,
                    _lib.ptr(grid), _lib.ptr(self.oob_total), _lib.stream_ptr(dev)))   # checked once per sequence
            return grid
        xs, ys, ts, ps = self.get_events(idx0, idx1)
        ts = (ts - ts[0]).astype(np.float32)
        return self.get_voxel_grid(torch.from_numpy(xs), torch.from_numpy(ys), torch.from_numpy(ts),
                                   torch.from_numpy(ps.astype(np.float32)))

    def check_bounds(self):
        """Raise IndexError (like the reference's index_put_) if any event fell outside the sensor."""
        if self.oob_total is not None and int(self.oob_total.item()) != 0:
            raise IndexError("%d events are out of bounds for sensor_resolution %s"
                             % (int(self.oob_total.item()), tuple(self.sensor_resolution)))

    def get_voxel_grid(self, xs, ys, ts, ps):
        return events_to_voxel_torch(xs, ys, ts, ps, self.num_bins, device=self._device(),
                                     sensor_size=self.sensor_resolution)

    def get_frame(self, index):
        return self.filehandle['images'][index][:, :, 0]

    def get_frame_tensor(self, index):
        """frame / 255 as float32 [1,H,W] on the device (dataset.py:84)."""
        dev = self._device()
        if self.resident:
            self._upload()
            src = self._dev_images[index]
        else:
            src = torch.from_numpy(np.ascontiguousarray(self.get_frame(index))).to(dev, non_blocking=True)
        out = torch.empty((1,) + tuple(src.shape), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().evk_u8_to_f32(_lib.ptr(src), _lib.ptr(out), src.numel(), _lib.stream_ptr(dev)))
        return out

    def get_events(self, idx0, idx1):
        xy = self.filehandle["xy"][idx0:idx1]
        xs = xy[:, 0].astype(np.float32)
        ys = xy[:, 1].astype(np.float32)
        ts = self.filehandle["t"][idx0:idx1]
        ps = self.filehandle["p"][idx0:idx1] * 2.0 - 1.0
        return xs, ys, ts, ps

    # ------------------------------------------------------------------ loading
    def load_data(self, data_path):
        if isinstance(data_path, dict):
            # in-memory sequence in the on-disk layout (synthetic benchmark streams): same keys as the .npy files
            a = data_path
            data = {}
            if all(k in a for k in ('images_ts', 'images', 'image_event_indices')):
                data["frame_stamps"] = np.asarray(a['images_ts'])
                data["images"] = a['images']
                data["image_event_indices"] = np.asarray(a['image_event_indices'])
                self.has_images = True
            else:
                self.has_images = False
            data["t"] = np.asarray(a['events_ts']).squeeze()
            data["xy"] = np.asarray(a['events_xy']).squeeze()
            data["p"] = np.asarray(a['events_p']).squeeze()
            data['path'] = '<memory>'
            if self.sensor_resolution is None and 'sensor_resolution' in a:
                self.sensor_resolution = list(a['sensor_resolution'])
            return self._finish_load(data, None)
        assert os.path.isdir(data_path), f'{data_path} is not a valid data_path'
        data = {}
        p = lambda name: os.path.join(data_path, name)
        if all(os.path.exists(p(f)) for f in ('images_ts.npy', 'images.npy', 'image_event_indices.npy')):
            data["frame_stamps"] = np.load(p('images_ts.npy'))
            data["images"] = np.load(p('images.npy'), mmap_mode='r')
            data["image_event_indices"] = np.load(p('image_event_indices.npy'))
            self.has_images = True
        else:
            self.has_images = False
        data["t"] = np.load(p('events_ts.npy'), mmap_mode='r').squeeze()
        data["xy"] = np.load(p('events_xy.npy'), mmap_mode='r').squeeze()
        data["p"] = np.load(p('events_p.npy'), mmap_mode='r').squeeze()
        data['path'] = data_path
        return self._finish_load(data, p("metadata.json"))

    def _finish_load(self, data, metadata_path):
        assert (len(data['p']) == len(data['xy']) and len(data['p']) == len(data['t'])), \
            "Number of events, timestamps and coordinates do not match"
        self.t0, self.tk = data['t'][0], data['t'][-1]
        self.num_events = len(data['p'])
        self.frame_ts = []
        if self.has_images:
            self.num_frames = len(data['images'])
            for ts in data["frame_stamps"]:
                self.frame_ts.append(ts.item())
            data["index"] = self.frame_ts
        else:
            self.num_frames = 0
        assert (len(self.frame_ts) == self.num_frames), "Number of frames and timestamps do not match"
        self.filehandle = data
        if self.sensor_resolution is None:
            if metadata_path is not None and os.path.exists(metadata_path):
                self.sensor_resolution = read_json(metadata_path)["sensor_resolution"]
            elif self.has_images and self.num_frames > 0:
                self.sensor_resolution = self.filehandle["images"][0].shape[:2]
            else:
                self.sensor_resolution = [np.max(self.filehandle["xy"][:, 1]) + 1,
                                          np.max(self.filehandle["xy"][:, 0]) + 1]

    def find_ts_index(self, timestamp):
        return np.searchsorted(self.filehandle["t"], timestamp)
