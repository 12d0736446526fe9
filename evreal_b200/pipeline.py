"""Lock-step evaluation of B independent sequences on one GPU: the throughput form of EVREAL's per-frame loop.

The reference evaluates one sequence at a time with batch 1 (eval.py:189-246, DataLoader defaults at
eval.py:72).  Sequences are independent (state reset per sequence, eval.py:197) and frames inside one are
strictly serial, so the B200 form of the loop runs frame ``i`` of B sequences together: ONE voxelizer launch for
the B windows, then ONE batched normalise+pad, network forward, crop, percentile normalisation and fused MSE/SSIM
(and LPIPS) launch for all B.  Per-sample arithmetic is unchanged (normalize_event_tensor statistics are per sample), so
every sequence gets the frames and scores it would get alone (tests/test_gpu_pipeline.py).

Stages 1 and 3 are frame-independent: with ``overlap=True`` the voxelizer / normaliser of step i+1 runs on a
"pre" stream and crop / percentile / metrics of step i-1 on a "post" stream while the network of step i owns the main
stream.  The network's input and output buffers are double-buffered inside the model handle for exactly this
(``evk_model_io_buffers``); the convolution kernels are persistent and fill every SM, so the side kernels run in the
tails between them instead of extending the critical path.

Two input modes:
  * ``resident=True``  -- raw event arrays (int16 xy, float64 t, uint8 p) and reference frames are uploaded once
    and every window is voxelized from HBM (bench.py ``value``);
  * ``resident=False`` -- the arrays stay in pinned HOST memory; every step's windows (13 B/event) and reference
    frames are copied host->device and scores + reconstructed frames are read back (bench.py ``e2e``).  The copies
    run on their own streams: the windows of step i+1 are staged (double-buffered) while step i computes, and the
    results of step i travel back while step i+1 runs, so PCIe time hides behind the network.
All device buffers are allocated once; a step allocates nothing.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .util import CropParameters


class SequenceBatch:
    RING = 4          # result slots (device + pinned host) in flight

    def __init__(self, model, datasets, event_tensor_normalization=False, post_process_norm='none', resident=True,
                 device=None, compute_metrics=True, offsets=None, counts=None, lpips=None, overlap=True, log_scores=False):
        """``offsets`` / ``counts`` (per sequence): step k processes item offsets[b] + k of sequence b for k < counts[b] and an
        empty window afterwards (states are per sample, so a finished sequence does not disturb the others); the batch then
        has max(counts) steps.  Default: item k of every sequence, min(len) steps.
        ``lpips``: None, or (variant name 'lpips' | 'lpips-vgg', state_dict) -- LPIPS of every (frame, reference) pair of a
        step in one batched call (the reference's queue of 4, utils/eval_metrics.py:141-148, is a per-pair function).
        ``log_scores``: keep every step's scores on the device in ``scores_log [steps, B, 3]`` (mse, ssim, lpips)."""
        _lib.require_cuda()
        self.lib = _lib.load()
        self.model = model
        self.datasets = list(datasets)
        self.B = len(self.datasets)
        self.normalize = bool(event_tensor_normalization)
        self.post = post_process_norm
        if self.post not in ('none', 'robust', 'standard', 'exprobust'):
            raise ValueError(f"Unrecognized normalization argument: {self.post}")
        self.resident = resident
        self.compute_metrics = compute_metrics
        self.overlap = bool(overlap)
        self.dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        ds0 = self.datasets[0]
        self.H, self.W = int(ds0.sensor_resolution[0]), int(ds0.sensor_resolution[1])
        self.bins = ds0.num_bins
        for ds in self.datasets:
            assert (int(ds.sensor_resolution[0]), int(ds.sensor_resolution[1])) == (self.H, self.W), \
                "sequences batched together must share the sensor resolution"
            assert ds.num_bins == self.bins, "sequences batched together must share num_bins"
            assert ds.has_images or not compute_metrics
        self.crop = CropParameters(self.W, self.H, model.num_encoders)
        self.Hp, self.Wp = self.crop.height_crop_size, self.crop.width_crop_size
        B, H, W, dev = self.B, self.H, self.W, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            model.to(dev)
            model._ensure(B, self.Hp, self.Wp, dev)            # device program + double-buffered input / output
        self.voxel = torch.zeros((B, self.bins, H, W), **f32)
        self.recon = torch.empty((B, 1, H, W), **f32)
        self.ref2 = torch.zeros((2, B, H, W), **f32)            # reference frames of two consecutive steps
        self.ref = self.ref2[0]
        # results rotate through RING device slots so that the device->host copy of step i (own stream) never races
        # the kernels of step i+1
        self.image_ring = torch.empty((self.RING, B, 1, H, W), **f32)
        self.scores_ring = torch.zeros((self.RING, B, 2), dtype=torch.float64, device=dev)
        self.lpips_ring = torch.zeros((self.RING, B), dtype=torch.float64, device=dev)
        self.scores = self.scores_ring[0]
        self.lpips_scores = self.lpips_ring[0]
        self.image = self.image_ring[0]
        self.oob_total = torch.zeros(1, dtype=torch.int32, device=dev)
        self.launches = 0            # kernels launched by the last step (bench.py gpu_launches)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._nstep = 0
        # per item of every sequence, once (dataset.py:33-102 via MemMapDataset.item_meta): event range, reference frame,
        # voxel / frame timestamps (the score gates of utils/eval_metrics.py:256-262 are evaluated from these on the host)
        self._offsets = [0] * B if offsets is None else [int(o) for o in offsets]
        self._counts = None if counts is None else [int(c) for c in counts]
        n_items = len(self)
        self._win = np.zeros((B, n_items, 3), dtype=np.int64)
        self.voxel_ts = np.zeros((B, n_items), dtype=np.float64)
        self.frame_ts = np.zeros((B, n_items), dtype=np.float64)
        max_win = 1
        for b, ds in enumerate(self.datasets):
            for i in range(n_items if self._counts is None else min(n_items, self._counts[b])):
                i0, i1, fi, vts, fts, _, _ = ds.item_meta(self._offsets[b] + i)
                self._win[b, i] = (i0, max(i1, i0), fi)
                self.voxel_ts[b, i], self.frame_ts[b, i] = vts, fts
            max_win = max(max_win, int((self._win[b, :, 1] - self._win[b, :, 0]).max()) if n_items else 1)
        self.max_win = max_win
        self._src = []
        self._base = np.zeros((B, 4), dtype=np.int64)      # raw base addresses (xy, t, p, images) of every sequence
        self._stream_up = None
        if resident:
            # threaded, chunked upload of every sequence (13 B/event + frames, _upload.py), issued in TIME SLICES: the events /
            # frames of the first steps of every sequence first, so the loop starts while later slices are still in flight
            from . import _upload
            up = _upload.get(dev)
            n_slices = max(1, min(8, n_items // 16))
            self._slice_of = [min(i * n_slices // max(n_items, 1), n_slices - 1) for i in range(max(n_items, 1))]
            slices = [[] for _ in range(n_slices)]
            for b, ds in enumerate(self.datasets):
                fh = ds.filehandle
                xy_h = np.ascontiguousarray(np.asarray(fh["xy"]) if np.asarray(fh["xy"]).dtype == np.int16 else np.asarray(fh["xy"]).astype(np.int16))
                t_h = np.ascontiguousarray(np.asarray(fh["t"], dtype=np.float64))
                p_h = np.asarray(fh["p"])
                p_h = np.ascontiguousarray(p_h if p_h.dtype == np.uint8 else p_h.astype(np.uint8))
                im_h = np.ascontiguousarray(np.asarray(fh["images"])[..., 0]) if ds.has_images else None
                xy, t, p = up.alloc_like(xy_h), up.alloc_like(t_h), up.alloc_like(p_h)
                im = up.alloc_like(im_h) if im_h is not None else None
                self._src.append((xy, t, p, im))
                self._base[b] = (xy.data_ptr(), t.data_ptr(), p.data_ptr(), im.data_ptr() if im is not None else 0)
                # slice j needs every event below the largest window end and every frame up to the largest frame index of its items
                e_lo = f_lo = 0
                for j in range(n_slices):
                    items = [i for i in range(n_items) if self._slice_of[i] == j]
                    last = j == n_slices - 1
                    e_hi = len(t_h) if last else max([int(self._win[b, i, 1]) for i in items] + [e_lo])
                    f_hi = (len(im_h) if im_h is not None else 0) if last else max([int(self._win[b, i, 2]) + 1 for i in items] + [f_lo])
                    e_hi, f_hi = max(e_hi, e_lo), max(f_hi, f_lo)
                    slices[j] += up.jobs(xy_h, xy, e_lo, e_hi) + up.jobs(t_h, t, e_lo, e_hi) + up.jobs(p_h, p, e_lo, e_hi)
                    if im_h is not None:
                        slices[j] += up.jobs(im_h, im, f_lo, f_hi)
                    e_lo, f_lo = e_hi, f_hi
            self._stream_up = _upload.StreamedUpload(up, slices)
        else:
            for b, ds in enumerate(self.datasets):
                fh = ds.filehandle
                xy = torch.from_numpy(np.ascontiguousarray(fh["xy"], dtype=np.int16)).pin_memory()
                t = torch.from_numpy(np.ascontiguousarray(fh["t"], dtype=np.float64)).pin_memory()
                p = torch.from_numpy(np.ascontiguousarray(fh["p"]).astype(np.uint8)).pin_memory()
                im = torch.from_numpy(np.ascontiguousarray(fh["images"][..., 0])).pin_memory() if ds.has_images else None
                self._src.append((xy, t, p, im))
                self._base[b] = (xy.data_ptr(), t.data_ptr(), p.data_ptr(), im.data_ptr() if im is not None else 0)
        self._windows = (_lib.EventWindow * B)()
        self._frames = (ctypes.c_void_p * B)()
        # numpy views of the two tables the batched kernels read (filled per step with four vector assignments instead of a Python
        # loop over the sequences: the lock-step loop of a small network is bound by this host code)
        self._windows_np = np.frombuffer(self._windows, dtype=np.int64).reshape(B, 4)      # columns: xy, t, pol, n
        self._frames_np = np.frombuffer(self._frames, dtype=np.int64)
        self.lpips_net = None
        if lpips is not None and compute_metrics:
            from .lpips import LpipsNet
            name, sd = lpips
            with torch.cuda.device(dev):
                self.lpips_net = LpipsNet(name, sd, H, W, batch=B)
        self.scores_log = None
        if log_scores:
            self.scores_log = torch.zeros((max(n_items, 1), B, 3), dtype=torch.float64, device=dev)
        # streams and the events that order the three stages of neighbouring steps
        if self.overlap:
            self.pre_stream = torch.cuda.Stream(dev)
            self.post_stream = torch.cuda.Stream(dev)
        else:
            self.pre_stream = self.post_stream = None
        self._pre_issued = None          # item whose voxel tensor is (being) written into the model's next input buffer
        self._pre_in_ptr = None
        self._pre_done = torch.cuda.Event()
        self._fwd_done = [torch.cuda.Event(), torch.cuda.Event()]      # by step parity
        self._post_done = [torch.cuda.Event(), torch.cuda.Event()]
        self._fwd_recorded = [False, False]
        self._post_recorded = [False, False]
        self.result_event = None
        if not resident:
            # double-buffered device staging: slot s is filled by the copy stream while slot s^1 is being voxelized
            self.st_xy = torch.empty((2, B, max_win, 2), dtype=torch.int16, device=dev)
            self.st_t = torch.empty((2, B, max_win), dtype=torch.float64, device=dev)
            self.st_p = torch.empty((2, B, max_win), dtype=torch.uint8, device=dev)
            self.st_im = torch.empty((2, B, H, W), dtype=torch.uint8, device=dev)
            self.host_scores = torch.empty((self.RING, B, 2), dtype=torch.float64).pin_memory()
            self.host_lpips = torch.empty((self.RING, B), dtype=torch.float64).pin_memory()
            self.host_image = torch.empty((self.RING, B, 1, H, W), dtype=torch.float32).pin_memory()
            self.copy_stream = torch.cuda.Stream(dev)
            self.d2h_stream = torch.cuda.Stream(dev)
            self._h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [None, None]                 # recorded on the pre stream after a slot was voxelized
            self._d2h_done = [None] * self.RING
            self._staged = [None, None]                   # item index held (or in flight) in each staging slot
            self._last_slot = 1
            self._host_windows = (_lib.EventWindow * B)()
            self._host_frames = (ctypes.c_void_p * B)()
            self._host_windows_np = np.frombuffer(self._host_windows, dtype=np.int64).reshape(B, 4)
            self._host_frames_np = np.frombuffer(self._host_frames, dtype=np.int64)
            self._b_index = np.arange(B, dtype=np.int64)
        if self._stream_up is None:
            torch.cuda.synchronize(dev)

    def __len__(self):
        if self._counts is not None:
            return max(self._counts) if self._counts else 0
        return min(len(ds) - o for ds, o in zip(self.datasets, self._offsets))

    def reset(self):
        self._sync_streams()                 # (not the background upload: the first steps run while later slices arrive)
        self.model.reset_states()            # (also rewinds the parity of the model's input / output buffers)
        self.oob_total.zero_()
        self._pre_issued = None
        self._nstep = 0
        self._fwd_recorded = [False, False]
        self._post_recorded = [False, False]

    # ------------------------------------------------------------------ host-mode staging
    def _stage(self, idx, slot):
        """Issue the host->device copies of item ``idx`` into staging slot ``slot`` on the copy stream."""
        B, lib = self.B, self.lib
        if self._consumed[slot] is not None:
            self.copy_stream.wait_event(self._consumed[slot])      # the kernels that read this slot have finished
        win = self._win[:, idx]
        hw, hn = self._host_windows, self._host_windows_np
        hn[:, 0] = self._base[:, 0] + win[:, 0] * 4
        hn[:, 1] = self._base[:, 1] + win[:, 0] * 8
        hn[:, 2] = self._base[:, 2] + win[:, 0]
        hn[:, 3] = win[:, 1] - win[:, 0]
        cs = ctypes.c_void_p(self.copy_stream.cuda_stream)
        _lib.check(lib.evk_stage_windows_h2d(hw, B, _lib.ptr(self.st_xy[slot]), _lib.ptr(self.st_t[slot]),
                                             _lib.ptr(self.st_p[slot]), self.max_win, cs))
        if self.compute_metrics:
            hf = self._host_frames
            self._host_frames_np[:] = self._base[:, 3] + win[:, 2] * (self.H * self.W)
            _lib.check(lib.evk_stage_frames_h2d(hf, B, self.H * self.W, _lib.ptr(self.st_im[slot]), cs))
        self._h2d_done[slot].record(self.copy_stream)
        self._staged[slot] = idx

    # ------------------------------------------------------------------ stage 1 of item idx -> the model's next input buffer
    def _issue_pre(self, idx, main, par):
        """Voxelize + normalise + pad item ``idx`` of every sequence into the input buffer of the model's next forward, and
        convert its reference frames into ref2[par]; on the pre stream when overlapping.  Returns kernel launches issued."""
        lib, B = self.lib, self.B
        pre = self.pre_stream if self.overlap else main
        st = ctypes.c_void_p(pre.cuda_stream)
        win = self._win[:, idx]
        n_events = int((win[:, 1] - win[:, 0]).sum())
        ws, fr = self._windows, self._frames
        launches = 0
        if self.overlap:
            # the buffers written below were last read by the forward / the metrics of the step before the previous one
            if self._fwd_recorded[par]:
                pre.wait_event(self._fwd_done[par])
            if self._post_recorded[par]:
                pre.wait_event(self._post_done[par])
        if self.resident:
            if self._stream_up is not None:
                self._stream_up.wait(self._slice_of[min(idx, len(self._slice_of) - 1)], pre)
            wn = self._windows_np
            wn[:, 0] = self._base[:, 0] + win[:, 0] * 4
            wn[:, 1] = self._base[:, 1] + win[:, 0] * 8
            wn[:, 2] = self._base[:, 2] + win[:, 0]
            wn[:, 3] = win[:, 1] - win[:, 0]
            self._frames_np[:] = self._base[:, 3] + win[:, 2] * (self.H * self.W)
        else:
            slot = 0 if self._staged[0] == idx else (1 if self._staged[1] == idx else None)
            if slot is None:                      # first step / non-sequential access: stage now
                slot = self._last_slot ^ 1
                self._stage(idx, slot)
            self._last_slot = slot
            pre.wait_event(self._h2d_done[slot])
            xy0, t0, p0, im0 = (self.st_xy[slot].data_ptr(), self.st_t[slot].data_ptr(), self.st_p[slot].data_ptr(),
                                self.st_im[slot].data_ptr())
            wn, bi = self._windows_np, self._b_index
            wn[:, 0] = xy0 + bi * (self.max_win * 4)
            wn[:, 1] = t0 + bi * (self.max_win * 8)
            wn[:, 2] = p0 + bi * self.max_win
            wn[:, 3] = win[:, 1] - win[:, 0]
            self._frames_np[:] = im0 + bi * (self.H * self.W)
            self._h2d_step = n_events * 13 + (B * self.H * self.W if self.compute_metrics else 0)
        in_ptr, _ = self.model.io_buffers()
        # empty windows -> zeros grid (dataset.py:59-71) is part of the batched call
        _lib.check(lib.evk_voxelize_raw_batch(ws, B, self.bins, self.H, self.W, _lib.ptr(self.voxel),
                                              _lib.ptr(self.oob_total), st))
        # grid clear + scatter (two reductions per event straight into the planar grid) for windows below ~2*groups*H*W events,
        # else scratch clear + scatter (one vector reduction per event) + gather: voxelize.cu
        direct = int((win[:, 1] - win[:, 0]).max()) <= 2 * ((self.bins - 1) // 3 + 1) * self.H * self.W
        launches += (2 if direct else 3) if n_events > 0 else 1
        if self.compute_metrics:
            _lib.check(lib.evk_u8_to_f32_batch(fr, B, self.H * self.W, _lib.ptr(self.ref2[par]), st))
            launches += 1
        if not self.resident:
            if self._consumed[slot] is None:
                self._consumed[slot] = torch.cuda.Event()
            self._consumed[slot].record(pre)
        _lib.check(lib.evk_normalize_pad(_lib.ptr(self.voxel), ctypes.c_void_p(in_ptr), B, self.bins, self.H, self.W,
                                         self.Hp, self.Wp, int(self.normalize), st))
        launches += 2 if self.normalize else 1
        self._pre_done.record(pre)
        self._pre_issued, self._pre_in_ptr, self._pre_events = idx, in_ptr, n_events
        return launches

    def step(self, idx, next_idx=None, sync=True):
        """Frame ``idx`` of every sequence.  Returns (scores [B,2] float64 (mse, ssim), image [B,1,H,W], n_events).
        The returned tensors are written by the post stream (host mode: pinned host tensors written by the copy stream): they
        are valid after ``self.result_event``.  ``sync=True`` (default) makes the current stream wait for it, so device
        tensors can be used right away; throughput loops pass ``sync=False`` (the next network forward then does not wait
        for this step's metrics) and read ``scores_log`` / call ``finish()`` at the end.  Host-mode results additionally
        need ``result_event.synchronize()`` or ``finish()`` before the host reads them.
        ``next_idx`` (default idx + 1) is the item whose windows are staged / voxelized while this one runs the network."""
        lib, dev, B = self.lib, self.dev, self.B
        main = torch.cuda.current_stream(dev)
        launches = 0
        h2d = d2h = 0
        ring = self._nstep % self.RING
        par = self._nstep & 1
        self._nstep += 1
        with torch.cuda.device(dev):
            self.model._ensure(B, self.Hp, self.Wp, dev)
            # ---- stage 1 (normally already issued by the previous step)
            if self._pre_issued != idx:
                launches += self._issue_pre(idx, main, par)
            n_events = self._pre_events
            if not self.resident:
                h2d += self._h2d_step
            in_ptr, out_ptr = self.model.io_buffers()
            assert in_ptr == self._pre_in_ptr, "model parity changed under the pipeline (call reset() after model.reset_states())"
            # ---- stage 2: the network, on the caller's stream
            if self.overlap:
                main.wait_event(self._pre_done)
                if self._post_recorded[par]:
                    main.wait_event(self._post_done[par])        # the output buffer of this parity has been consumed
            self.model.forward_raw(in_ptr, out_ptr)
            launches += self.model.last_launch_count()
            self._fwd_done[par].record(main)
            self._fwd_recorded[par] = True
            # ---- stage 1 of the next item, behind the network just enqueued
            nxt = idx + 1 if next_idx is None else next_idx
            if 0 <= nxt < self._win.shape[1] and nxt != idx:
                if not self.resident and self._staged[0] != nxt and self._staged[1] != nxt:
                    self._stage(nxt, self._last_slot ^ 1)
                launches_next = self._issue_pre(nxt, main, par ^ 1)
            else:
                launches_next = 0
                self._pre_issued = None
            # ---- stage 3 on the post stream
            post = self.post_stream if self.overlap else main
            st = ctypes.c_void_p(post.cuda_stream)
            if self.overlap:
                post.wait_event(self._fwd_done[par])
            if not self.resident and self._d2h_done[ring] is not None:
                post.wait_event(self._d2h_done[ring])             # slot's previous results have left the device
            image = self.image_ring[ring]
            row = self.scores_log[idx] if self.scores_log is not None and idx < self.scores_log.shape[0] else None
            scores = self.scores_ring[ring]
            lp = self.lpips_ring[ring]
            # the ring slot is what travels to the host: without a percentile pass the crop writes it directly (a shared
            # crop buffer would be overwritten by the next step while the device->host copy of this one is still reading)
            cropped = self.recon if self.post != 'none' else image
            _lib.check(lib.evk_crop(ctypes.c_void_p(out_ptr), _lib.ptr(cropped), B, 1, self.Hp, self.Wp, self.H, self.W, st))
            launches += 1
            if self.post != 'none':
                q = (0.0, 100.0) if self.post == 'standard' else (1.0, 99.0)
                _lib.check(lib.evk_percentile_normalize(_lib.ptr(self.recon), _lib.ptr(image), B, self.H * self.W,
                                                        q[0], q[1], int(self.post == 'exprobust'), st))
                launches += 1
            if self.compute_metrics:
                ref = self.ref2[par]
                # clip of EvalMetricsTracker.update (utils/eval_metrics.py:253-255) fused into the metric kernels
                _lib.check(lib.evk_mse_ssim(_lib.ptr(image), _lib.ptr(ref), B, self.H, self.W, 1, _lib.ptr(scores), st))
                launches += 2
                if self.lpips_net is not None:
                    _lib.check(lib.evk_lpips_forward(self.lpips_net.handle, _lib.ptr(image), _lib.ptr(ref), B, _lib.ptr(lp), st))
                    launches += self.lpips_net.launches_per_forward
                if row is not None:
                    with torch.cuda.stream(post):
                        row[:, :2].copy_(scores, non_blocking=True)
                        if self.lpips_net is not None:
                            row[:, 2].copy_(lp, non_blocking=True)
                self.ref = ref
            self._post_done[par].record(post)
            self._post_recorded[par] = True
            self.result_event = self._post_done[par]
            self.scores, self.image, self.lpips_scores = scores, image, lp
            if not self.resident:
                self.d2h_stream.wait_event(self._post_done[par])
                with torch.cuda.stream(self.d2h_stream):
                    self.host_scores[ring].copy_(scores, non_blocking=True)
                    self.host_image[ring].copy_(image, non_blocking=True)
                    if self.lpips_net is not None:
                        self.host_lpips[ring].copy_(lp, non_blocking=True)
                if self._d2h_done[ring] is None:
                    self._d2h_done[ring] = torch.cuda.Event()
                self._d2h_done[ring].record(self.d2h_stream)
                self.result_event = self._d2h_done[ring]
                d2h += scores.numel() * 8 + image.numel() * 4 + (lp.numel() * 8 if self.lpips_net is not None else 0)
                scores, image = self.host_scores[ring], self.host_image[ring]
                self.host_lpips_scores = self.host_lpips[ring]
        self.launches, self.h2d_bytes, self.d2h_bytes = launches + launches_next, h2d, d2h
        if sync and self.overlap:
            main.wait_event(self._post_done[par])
        return scores, image, n_events

    def wait_results(self):
        """Make the current stream wait for the results of the last step (device-side; does not block the host)."""
        if self.result_event is not None:
            torch.cuda.current_stream(self.dev).wait_event(self.result_event)

    def finish(self):
        """Wait (on the host) for every stage and copy in flight, the background upload included."""
        if self._stream_up is not None:
            self._stream_up.finish()
        self._sync_streams()

    def wait_uploaded(self):
        """Block until every sequence is completely resident in HBM (benchmarks: nothing may still be crossing PCIe when a
        timed region that claims resident inputs starts)."""
        if self._stream_up is not None:
            self._stream_up.finish()

    def _sync_streams(self):
        if self.overlap:
            self.pre_stream.synchronize()
            self.post_stream.synchronize()
        if not self.resident:
            self.copy_stream.synchronize()
            self.d2h_stream.synchronize()

    def join(self, stream=None):
        """Make ``stream`` (default: the current stream) wait for all side-stream work issued so far -- used to bracket a
        timed region with events recorded on one stream."""
        s = stream if stream is not None else torch.cuda.current_stream(self.dev)
        for side in (self.pre_stream, self.post_stream, getattr(self, 'copy_stream', None), getattr(self, 'd2h_stream', None)):
            if side is not None:
                e = torch.cuda.Event()
                e.record(side)
                s.wait_event(e)

    def check_bounds(self):
        n = int(self.oob_total.item())
        if n != 0:
            raise IndexError("%d events are out of bounds for sensor_resolution %s" % (n, (self.H, self.W)))
