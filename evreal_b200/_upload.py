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
    def __init__(self, device, workers=None, chunk_bytes=8 << 20):
        import os
        if workers is None:
            world = max(int(os.environ.get("WORLD_SIZE", "1")), 1)
            workers = max(2, min(8, (os.cpu_count() or 4) // world))
            if os.environ.get("EVK_UPLOAD_WORKERS"):
                workers = max(1, int(os.environ["EVK_UPLOAD_WORKERS"]))
        self.dev = torch.device(device)
        self.workers, self.chunk = workers, chunk_bytes
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pinned = [[torch.empty(chunk_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(workers)]
        self.views = [[b.numpy() for b in pair] for pair in self.pinned]
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(workers)]
        self.events = [[None, None] for _ in range(workers)]
        self.locks = [threading.Lock() for _ in range(workers)]      # a worker slot (pinned pair + stream) serves one upload at a time

    def _work(self, j, jobs):
        """jobs: [(src uint8 numpy view, dst uint8 device tensor view)] of at most chunk bytes each."""
        stream = self.streams[j]
        with self.locks[j]:
            self._copy_jobs(j, stream, jobs)

    def _copy_jobs(self, j, stream, jobs):
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

    def alloc_like(self, a):
        return torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype, device=self.dev)

    def jobs(self, a, t, lo=0, hi=None):
        """Copy jobs (chunks) for rows [lo, hi) of the C-contiguous host array ``a`` into the same rows of device tensor ``t``."""
        hi = a.shape[0] if hi is None else hi
        if hi <= lo:
            return []
        src = a[lo:hi].reshape(-1).view(np.uint8)
        dst = t[lo:hi].reshape(-1).view(torch.uint8)
        return [(src[off:off + self.chunk], dst[off:off + self.chunk]) for off in range(0, src.shape[0], self.chunk)]

    def upload(self, arrays):
        """[numpy array, ...] (C-contiguous) -> [device tensor, ...] of the same dtypes / shapes; returns after all copies
        have completed."""
        outs, jobs = [], []
        for a in arrays:
            a = np.ascontiguousarray(a)
            t = self.alloc_like(a)
            outs.append(t)
            jobs += self.jobs(a, t)
        if not jobs:
            return outs
        with torch.cuda.device(self.dev):
            futs = [self.pool.submit(self._work, j, jobs[j::self.workers]) for j in range(self.workers)]
            for f in futs:
                f.result()
            for s in self.streams:
                s.synchronize()
        return outs


class StreamedUpload:
    """Handle of an upload running in the background, slice by slice (a slice = the events / frames of a range of steps of
    every sequence of a lock-step batch): ``wait(j, stream)`` blocks the HOST until slice j's copies have been issued and
    makes ``stream`` wait (on the device) until they have completed, so the first steps run while later slices are still
    crossing PCIe."""

    def __init__(self, uploader, slices):
        self.up = uploader
        self.n = len(slices)
        W = uploader.workers
        self.issued = [threading.Event() for _ in range(self.n)]
        self.pending = [W] * self.n
        self.lock = threading.Lock()
        self.events = [[torch.cuda.Event() for _ in range(W)] for _ in range(self.n)]
        self.waited = -1
        self.error = None
        self.futs = [uploader.pool.submit(self._work, j, [sl[j::W] for sl in slices]) for j in range(W)]

    def _work(self, j, per_slice):
        up = self.up
        try:
            with up.locks[j], torch.cuda.device(up.dev):
                k = 0
                for si, jobs in enumerate(per_slice):
                    for src, dst in jobs:
                        slot = k & 1
                        k += 1
                        ev = up.events[j][slot]
                        if ev is not None:
                            ev.synchronize()
                        n = src.shape[0]
                        np.copyto(up.views[j][slot][:n], src)
                        with torch.cuda.stream(up.streams[j]):
                            dst.copy_(up.pinned[j][slot][:n], non_blocking=True)
                            if ev is None:
                                ev = up.events[j][slot] = torch.cuda.Event()
                            ev.record(up.streams[j])
                    self.events[si][j].record(up.streams[j])
                    with self.lock:
                        self.pending[si] -= 1
                        if self.pending[si] == 0:
                            self.issued[si].set()
        except Exception as e:                  # never leave the consumer waiting
            self.error = e
            for ev in self.issued:
                ev.set()

    def wait(self, j, stream):
        j = min(j, self.n - 1)
        while self.waited < j:
            self.waited += 1
            self.issued[self.waited].wait()
            if self.error is not None:
                raise self.error
            for ev in self.events[self.waited]:
                stream.wait_event(ev)

    def finish(self):
        for f in self.futs:
            f.result()
        if self.error is not None:
            raise self.error
        for s in self.up.streams:
            s.synchronize()


def get(device):
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    with _lock:
        if key not in _uploaders:
            _uploaders[key] = Uploader(torch.device('cuda', key[1]))
        return _uploaders[key]
