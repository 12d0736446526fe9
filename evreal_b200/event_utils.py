"""Host-side mirror of EVREAL's ``utils/event_utils.py`` on top of the CUDA voxelizer.

Same names, argument meaning and error behaviour as the reference
(``utils/event_utils.py:4-59``); inputs may live on the CPU (they are copied to
the GPU) or already on the GPU.  The result always lives on the CUDA device --
there is no CPU code path.
"""
import ctypes

import torch

from . import _lib


def _dev(device):
    _lib.require_cuda()
    if device is None or torch.device(device).type != "cuda":
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device(device)
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


def _f32(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def events_to_voxel_torch(xs, ys, ts, ps, num_bins, device=None, sensor_size=(180, 240), check_bounds=True):
    """Voxel grid [num_bins, H, W] with temporal bilinear interpolation (utils/event_utils.py:27-59).

    ``check_bounds`` (default on, like the reference's IndexError) costs one
    device->host read of a 4-byte counter; the streaming pipeline turns it off
    and checks once per sequence instead.
    """
    assert (len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps))
    device = _dev(device if device is not None else (xs.device if torch.is_tensor(xs) else None))
    lib = _lib.load()
    n = len(xs)
    if n == 0:
        raise IndexError("index -1 is out of bounds for dimension 0 with size 0")   # ts[-1] in the reference
    with torch.cuda.device(device):
        x, y, t, p = (_f32(v, device) for v in (xs, ys, ts, ps))
        H, W = int(sensor_size[0]), int(sensor_size[1])
        grid = torch.empty((num_bins, H, W), dtype=torch.float32, device=device)
        oob = torch.zeros(1, dtype=torch.int32, device=device) if check_bounds else None
        _lib.check(lib.evk_voxelize(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), n, num_bins, H, W,
                                    _lib.ptr(grid), _lib.ptr(oob) if check_bounds else None, _lib.stream_ptr(device)))
        if check_bounds and int(oob.item()) != 0:
            raise IndexError("index is out of bounds for sensor_size %s (%d events)" % ((H, W), int(oob.item())))
    return grid


def events_to_image_torch(xs, ys, ps, device=None, sensor_size=(180, 240)):
    """Scatter-add of weighted events into one H x W image (utils/event_utils.py:4-24):
    the single-bin case of the voxelizer with all events at the same normalised time."""
    # num_bins == 1 -> every event has t_norm == 0 and therefore full weight in bin 0
    zeros = torch.zeros(len(xs), dtype=torch.float32)
    return events_to_voxel_torch(xs, ys, zeros, ps, 1, device, sensor_size)[0]


def events_to_voxel_raw(xy, t, p, num_bins, device=None, sensor_size=(180, 240), check_bounds=True):
    """Raw on-disk window (int16 xy pairs, float64 absolute t, uint8 polarity) -> voxel grid.

    Fuses ``MemMapDataset.get_events`` / ``__getitem__`` casts (dataset.py:222-228, :52-58)
    into the kernel so the host never touches per-event data (13 B/event over PCIe
    instead of 16).
    """
    device = _dev(device)
    lib = _lib.load()
    n = len(t)
    if n == 0:
        raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
    with torch.cuda.device(device):
        xy = torch.as_tensor(xy).to(device=device, dtype=torch.int16, non_blocking=True).contiguous()
        t = torch.as_tensor(t).to(device=device, dtype=torch.float64, non_blocking=True).contiguous()
        p = torch.as_tensor(p).to(device=device, dtype=torch.uint8, non_blocking=True).contiguous()
        H, W = int(sensor_size[0]), int(sensor_size[1])
        grid = torch.empty((num_bins, H, W), dtype=torch.float32, device=device)
        oob = torch.zeros(1, dtype=torch.int32, device=device) if check_bounds else None
        _lib.check(lib.evk_voxelize_raw(_lib.ptr(xy), _lib.ptr(t), _lib.ptr(p), n, num_bins, H, W, _lib.ptr(grid),
                                        _lib.ptr(oob) if check_bounds else None, _lib.stream_ptr(device)))
        if check_bounds and int(oob.item()) != 0:
            raise IndexError("index is out of bounds for sensor_size %s" % ((H, W),))
    return grid
