"""Host -> HBM upload of whole sequences (the once-per-sequence ingest of `resident` mode; the reference's counterpart is
the DataLoader reading memmaps, eval.py:72 / dataset.py:222-250).

A plain ``torch.from_numpy(memmap).to(device)`` is one synchronous copy through the driver's small staging buffer
(a few GB/s from page-cache-backed memory), and pinning the whole sequence first is a second full pass over it.  Here the
array is cut into chunks that W worker threads copy into their own pair of pinned buffers (numpy's copy releases the GIL)
and hand to the copy engine on their own stream: page-cache reads, pinned stores and PCIe DMA of different chunks overlap.
One Uploader per (process, device); its pinned buffers are allocated once.
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_uploaders = {}
_lock = threading.Lock()


class Uploader:
    def __init__(self, device, workers=4, chunk_bytes=16 << 20):
        self.dev = torch.device(device)
        self.workers, self.chunk = workers, chunk_bytes
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pinned = [[torch.empty(chunk_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(workers)]
        self.views = [[b.numpy() for b in pair] for pair in self.pinned]
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(workers)]
        self.events = [[None, None] for _ in range(workers)]

    def _work(self, j, jobs):
        """jobs: [(src uint8 numpy view, dst uint8 device tensor view)] of at most chunk bytes each."""
        stream = self.streams[j]
        for k, (src, dst) in enumerate(jobs):
            slot = k & 1
            ev = self.events[j][slot]
            if ev is not None:
                ev.synchronize()                         # the DMA that last read this pinned buffer has finished
            n = src.shape[0]
            np.copyto(self.views[j][slot][:n], src)
            with torch.cuda.stream(stream):
                dst.copy_(self.pinned[j][slot][:n], non_blocking=True)
                if ev is None:
                    ev = self.events[j][slot] = torch.cuda.Event()
                ev.record(stream)

    def upload(self, arrays):
        """[numpy array, ...] (C-contiguous) -> [device tensor, ...] of the same dtypes / shapes; returns after all copies
        have completed."""
        outs, jobs = [], []
        for a in arrays:
            a = np.ascontiguousarray(a)
            t = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype, device=self.dev)
            outs.append(t)
            src = a.reshape(-1).view(np.uint8)
            dst = t.reshape(-1).view(torch.uint8)
            for off in range(0, src.shape[0], self.chunk):
                jobs.append((src[off:off + self.chunk], dst[off:off + self.chunk]))
        if not jobs:
            return outs
        with torch.cuda.device(self.dev):
            futs = [self.pool.submit(self._work, j, jobs[j::self.workers]) for j in range(self.workers)]
            for f in futs:
                f.result()
            for s in self.streams:
                s.synchronize()
        return outs


def get(device):
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    with _lock:
        if key not in _uploaders:
            _uploaders[key] = Uploader(torch.device('cuda', key[1]))
        return _uploaders[key]
