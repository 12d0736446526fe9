"""Host-side mirror of EVREAL's ``utils/eval_utils.py`` (+ ``eval.post_process_normalization``).

Numerical helpers run on the GPU through the C ABI; numpy inputs are accepted
for drop-in use (copied to the device and back), CUDA tensors stay resident.
"""
from os.path import join
from pathlib import Path

import numpy as np
import torch

from . import _lib


def ensure_dir(dirname):
    dirname = Path(dirname)
    if not dirname.is_dir():
        dirname.mkdir(parents=True, exist_ok=True)


def _to_cuda_f32(img):
    _lib.require_cuda()
    if isinstance(img, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).cuda(non_blocking=True), True
    return img.to(device='cuda' if not img.is_cuda else img.device, dtype=torch.float32).contiguous(), False


def percentile_normalize(img, q_min, q_max, apply_exp=False, batched=False):
    """(img - P_qmin) / (P_qmax - P_qmin), numpy-style linear-interpolated percentiles.
    ``batched``: treat dim 0 as independent images."""
    x, was_numpy = _to_cuda_f32(img)
    n = x.shape[0] if batched else 1
    numel = x.numel() // n
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().evk_percentile_normalize(_lib.ptr(x), _lib.ptr(out), n, numel, float(q_min), float(q_max),
                                                        int(bool(apply_exp)), _lib.stream_ptr(x.device)))
    return out.cpu().numpy() if was_numpy else out


def robust_min(img, p=5):
    return np.percentile(np.asarray(img).ravel(), p)


def robust_max(img, p=95):
    return np.percentile(np.asarray(img).ravel(), p)


def normalize(img, q_min=10, q_max=90):
    """utils/eval_utils.py:23-35 on the GPU."""
    return percentile_normalize(img, q_min, q_max)


def post_process_normalization(img, norm):
    """eval.py:380-395."""
    if norm == 'robust':
        return percentile_normalize(img, 1, 99)
    if norm == 'standard':
        return percentile_normalize(img, 0, 100)
    if norm == 'none':
        return img
    if norm == 'exprobust':
        return percentile_normalize(img, 1, 99, apply_exp=True)
    raise ValueError(f"Unrecognized normalization argument: {norm}")


def torch2cv2(image):
    """utils/eval_utils.py:38-43: tensor -> numpy on the host, channels last when there is a channel axis."""
    arr = image.squeeze().cpu().numpy()
    return arr.transpose(1, 2, 0) if arr.ndim == 3 else arr


def cv2torch(image, num_ch=1):
    """utils/eval_utils.py:46-54: [H,W] -> [1,num_ch,H,W] (the plane repeated), [C,H,W] -> [1,C,H,W]."""
    t = torch.as_tensor(image)
    if t.dim() == 2:
        t = t[None]
        if num_ch > 1:
            t = t.expand(num_ch, -1, -1).contiguous()
    return t[None] if t.dim() == 3 else t


class AsyncWriter:
    """Output files off the critical path: the reference appends a line to 2-4 text files and encodes one PNG
    per frame inside its loop (utils/eval_utils.py:57-84).  Here the loop only enqueues; ONE worker thread writes in order.

    ``append(path, text)`` -- text appended to ``path`` (same bytes, same order as the synchronous writers).
    ``png(path, frame, event)`` -- ``frame`` is a uint8 array or a pinned host tensor that becomes valid when the CUDA
    ``event`` (device->host copy on a side stream) has completed; the worker waits for it, never the loop.
    ``flush()`` blocks until everything queued so far is on disk (called once per sequence); errors raised in the worker
    surface there."""

    def __init__(self, max_pending=256):
        import queue
        import threading
        self._q = queue.Queue(maxsize=max_pending)
        self._err = None
        self._t = threading.Thread(target=self._run, name="evk-writer", daemon=True)
        self._t.start()

    def _run(self):
        while True:
            job = self._q.get()
            try:
                if job is None:
                    return
                kind, path, payload, event = job
                if self._err is not None:
                    continue
                if kind == 'append':
                    with open(path, 'a', encoding="utf-8") as f:
                        f.write(payload)
                elif kind == 'truncate':
                    open(path, 'w', encoding="utf-8").close()
                else:
                    import cv2
                    if event is not None:
                        event.synchronize()
                    arr = payload.numpy() if hasattr(payload, 'numpy') else payload
                    cv2.imwrite(path, arr)
            except Exception as e:          # surfaced by flush()
                self._err = e
            finally:
                self._q.task_done()

    def append(self, path, text):
        self._q.put(('append', path, text, None))

    def truncate(self, path):
        self._q.put(('truncate', path, None, None))

    def png(self, path, frame, event=None):
        self._q.put(('png', path, frame, event))

    def pending(self):
        return self._q.unfinished_tasks

    def flush(self):
        self._q.join()
        if self._err is not None:
            e, self._err = self._err, None
            raise e

    def close(self):
        self._q.put(None)
        self._t.join()


_writer = None          # module-wide writer installed by evaluate() (None = synchronous writes, like the reference)


def set_writer(writer):
    global _writer
    old, _writer = _writer, writer
    return old


def _append(path, rows, fmt):
    text = ''.join(fmt.format(k, v) for k, v in rows)
    if _writer is not None:
        _writer.append(path, text)
        return
    with open(path, 'a', encoding="utf-8") as f:
        f.write(text)


def truncate_file(path):
    if _writer is not None:
        _writer.truncate(path)
        return
    open(path, 'w', encoding="utf-8").close()


def append_timestamp(path, description, timestamp):
    """'<index> <timestamp with 15 decimals>' (utils/eval_utils.py:57-59)."""
    _append(path, [(description, timestamp)], '{} {:.15f}\n')


def append_result(path, description, result, is_int=False):
    """'<index> <score>' lines, 5 decimals unless is_int; a list of results pairs up with a list of indices
    (utils/eval_utils.py:62-77)."""
    rows = list(zip(description, result)) if isinstance(result, list) else [(description, result)]
    _append(path, rows, '{} {}\n' if is_int else '{} {:.5f}\n')


class _PngStaging:
    """Pinned host slots + a copy stream for frames on their way to the PNG writer."""
    SLOTS = 16

    def __init__(self):
        self.slots = {}
        self.events = {}
        self.next = 0
        self.stream = None


_png_staging = _PngStaging()


def save_inferred_image(folder, image, idx):
    """PNG writer (utils/eval_utils.py:80-84): frame_<idx>.png = uint8(round(image * 255)).  A CUDA frame is quantised on the
    device (evk_quantize_u8) and, when an AsyncWriter is installed, copied to a pinned slot on a side stream and encoded by
    the writer thread; numpy frames and the synchronous mode follow the reference literally."""
    png_path = join(folder, 'frame_{:010d}.png'.format(idx))
    if torch.is_tensor(image) and image.is_cuda and image.dim() == 2:
        dev = image.device
        img = image.float().contiguous()
        q = torch.empty(img.shape, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().evk_quantize_u8(_lib.ptr(img), _lib.ptr(q), img.numel(), _lib.stream_ptr(dev)))
        if _writer is None:
            import cv2
            cv2.imwrite(png_path, q.cpu().numpy())
            return
        st = _png_staging
        if st.stream is None:
            st.stream = torch.cuda.Stream(dev)
        key = (st.next % st.SLOTS, tuple(img.shape))
        st.next += 1
        if key in st.events:
            st.events[key].synchronize()              # the slot's previous frame has reached the host ...
            while _writer.pending() > st.SLOTS // 2:  # ... and (bounded queue) the writer is not a whole ring behind
                import time
                time.sleep(0.0005)
        else:
            st.slots[key] = torch.empty(img.shape, dtype=torch.uint8).pin_memory()
            st.events[key] = torch.cuda.Event()
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        st.stream.wait_event(done)
        with torch.cuda.stream(st.stream):
            st.slots[key].copy_(q, non_blocking=True)
            q.record_stream(st.stream)
        st.events[key].record(st.stream)
        # the writer encodes a private copy: the pinned slot is reused SLOTS frames later
        _writer.png(png_path, _SlotCopy(st.slots[key], st.events[key]), None)
        return
    arr = image.squeeze().cpu().numpy() if torch.is_tensor(image) else np.asarray(image)
    arr = np.round(arr * 255).astype(np.uint8)
    if _writer is not None:
        _writer.png(png_path, arr, None)
        return
    import cv2
    cv2.imwrite(png_path, arr)


class _SlotCopy:
    """A pinned slot + the event that makes it valid; ``numpy()`` (called by the writer thread) waits and copies out."""

    def __init__(self, slot, event):
        self.slot, self.event = slot, event

    def numpy(self):
        self.event.synchronize()
        return self.slot.numpy().copy()
