"""Lock-step evaluation of B independent sequences on one GPU: the throughput form of EVREAL's per-frame loop.

The reference evaluates one sequence at a time with batch 1 (eval.py:189-246, DataLoader defaults at
eval.py:72).  Sequences are independent (state reset per sequence, eval.py:197) and frames inside one are
strictly serial, so the B200 form of the loop runs frame ``i`` of B sequences together: one voxelizer launch
per window, then ONE batched normalise+pad, network forward, crop, percentile normalisation and fused MSE/SSIM
launch for all B.  Per-sample arithmetic is unchanged (normalize_event_tensor statistics are per sample), so
every sequence gets the frames and scores it would get alone (tests/test_gpu_pipeline.py).

Two input modes:
  * ``resident=True``  -- raw event arrays (int16 xy, float64 t, uint8 p) and reference frames are uploaded once
    and every window is voxelized from HBM (bench.py ``value``);
  * ``resident=False`` -- the arrays stay in pinned HOST memory; each step copies its windows (13 B/event) and
    reference frames host->device inside the step and reads scores + reconstructed frames back (bench.py ``e2e``).
All device buffers are allocated once; a step allocates nothing.
"""
import numpy as np
import torch

from . import _lib
from .util import CropParameters


class SequenceBatch:
    def __init__(self, model, datasets, event_tensor_normalization=False, post_process_norm='none', resident=True,
                 device=None, compute_metrics=True):
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
        self.dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        ds0 = self.datasets[0]
        self.H, self.W = int(ds0.sensor_resolution[0]), int(ds0.sensor_resolution[1])
        self.bins = ds0.num_bins
        for ds in self.datasets:
            assert (int(ds.sensor_resolution[0]), int(ds.sensor_resolution[1])) == (self.H, self.W), \
                "sequences batched together must share the sensor resolution"
            assert ds.has_images or not compute_metrics
        self.crop = CropParameters(self.W, self.H, model.num_encoders)
        self.Hp, self.Wp = self.crop.height_crop_size, self.crop.width_crop_size
        B, H, W, dev = self.B, self.H, self.W, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        self.voxel = torch.zeros((B, self.bins, H, W), **f32)
        self.padded = torch.empty((B, self.bins, self.Hp, self.Wp), **f32)
        self.recon_p = torch.empty((B, 1, self.Hp, self.Wp), **f32)
        self.recon = torch.empty((B, 1, H, W), **f32)
        self.image = torch.empty((B, 1, H, W), **f32)
        self.ref = torch.zeros((B, H, W), **f32)
        self.scores = torch.zeros((B, 2), dtype=torch.float64, device=dev)
        self.oob_total = torch.zeros(1, dtype=torch.int32, device=dev)
        self.launches = 0            # kernels launched by the last step (bench.py gpu_launches)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._src = []
        max_win = 1
        for ds in self.datasets:
            for i in range(len(ds)):
                i0, i1, _ = ds.window(i)
                max_win = max(max_win, int(i1) - int(i0))
        for ds in self.datasets:
            fh = ds.filehandle
            xy = torch.from_numpy(np.ascontiguousarray(fh["xy"], dtype=np.int16)).pin_memory()
            t = torch.from_numpy(np.ascontiguousarray(fh["t"], dtype=np.float64)).pin_memory()
            p = torch.from_numpy(np.ascontiguousarray(fh["p"]).astype(np.uint8)).pin_memory()
            im = torch.from_numpy(np.ascontiguousarray(fh["images"][..., 0])).pin_memory() if ds.has_images else None
            if resident:
                xy, t, p = (a.to(dev, non_blocking=True) for a in (xy, t, p))
                im = im.to(dev, non_blocking=True) if im is not None else None
            self._src.append((xy, t, p, im))
        if not resident:
            # per-stream staging (windows are copied host->device every step); t first so it stays 8-byte aligned
            self.st_xy = torch.empty((B, max_win, 2), dtype=torch.int16, device=dev)
            self.st_t = torch.empty((B, max_win), dtype=torch.float64, device=dev)
            self.st_p = torch.empty((B, max_win), dtype=torch.uint8, device=dev)
            self.st_im = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
            # results land in a ring of pinned host slots; a slot is reused only after its copy completed
            self.ring = 4
            self.host_scores = torch.empty((self.ring, B, 2), dtype=torch.float64).pin_memory()
            self.host_image = torch.empty((self.ring, B, 1, H, W), dtype=torch.float32).pin_memory()
            self._slot_done = [None] * self.ring
            self._nstep = 0
        torch.cuda.synchronize(dev)

    def __len__(self):
        return min(len(ds) for ds in self.datasets)

    def reset(self):
        self.model.reset_states()
        self.oob_total.zero_()

    def step(self, idx):
        """Frame ``idx`` of every sequence.  Returns (scores [B,2] float64 (mse, ssim), image [B,1,H,W], n_events).
        In host mode the returned tensors are pinned host tensors (valid after the returned event / a sync)."""
        lib, dev, B = self.lib, self.dev, self.B
        st = _lib.stream_ptr(dev)
        launches = 0
        h2d = d2h = 0
        n_events = 0
        with torch.cuda.device(dev):
            for b, ds in enumerate(self.datasets):
                i0, i1, frame_index = ds.window(idx)
                i0, i1 = int(i0), int(i1)
                n = max(i1 - i0, 0)
                n_events += n
                xy, t, p, im = self._src[b]
                if n > 0:
                    if self.resident:
                        wxy, wt, wp = xy[i0:i1], t[i0:i1], p[i0:i1]
                    else:
                        wxy, wt, wp = self.st_xy[b, :n], self.st_t[b, :n], self.st_p[b, :n]
                        wxy.copy_(xy[i0:i1], non_blocking=True)
                        wt.copy_(t[i0:i1], non_blocking=True)
                        wp.copy_(p[i0:i1], non_blocking=True)
                        h2d += n * 13
                    _lib.check(lib.evk_voxelize_raw(_lib.ptr(wxy), _lib.ptr(wt), _lib.ptr(wp), n, self.bins, self.H,
                                                    self.W, _lib.ptr(self.voxel[b]), _lib.ptr(self.oob_total), st))
                    launches += 1
                else:
                    self.voxel[b].zero_()          # empty window -> zeros grid (dataset.py:59-71)
                if self.compute_metrics:
                    if self.resident:
                        src = im[frame_index]
                    else:
                        src = self.st_im[b]
                        src.copy_(im[frame_index], non_blocking=True)
                        h2d += src.numel()
                    _lib.check(lib.evk_u8_to_f32(_lib.ptr(src), _lib.ptr(self.ref[b]), src.numel(), st))
                    launches += 1
            _lib.check(lib.evk_normalize_pad(_lib.ptr(self.voxel), _lib.ptr(self.padded), B, self.bins, self.H, self.W,
                                             self.Hp, self.Wp, int(self.normalize), st))
            launches += 2 if self.normalize else 1
            out = self.model.forward_into(self.padded, self.recon_p)
            launches += self.model.last_launch_count()
            _lib.check(lib.evk_crop(_lib.ptr(out), _lib.ptr(self.recon), B, 1, self.Hp, self.Wp, self.H, self.W, st))
            launches += 1
            image = self.recon
            if self.post != 'none':
                q = (0.0, 100.0) if self.post == 'standard' else (1.0, 99.0)
                _lib.check(lib.evk_percentile_normalize(_lib.ptr(self.recon), _lib.ptr(self.image), B, self.H * self.W,
                                                        q[0], q[1], int(self.post == 'exprobust'), st))
                launches += 1
                image = self.image
            if self.compute_metrics:
                # clip of EvalMetricsTracker.update (utils/eval_metrics.py:253-255) fused into the metric kernel
                _lib.check(lib.evk_mse_ssim(_lib.ptr(image), _lib.ptr(self.ref), B, self.H, self.W, 1,
                                            _lib.ptr(self.scores), st))
                launches += 2
            scores = self.scores
            if not self.resident:
                slot = self._nstep % self.ring
                self._nstep += 1
                if self._slot_done[slot] is not None:
                    self._slot_done[slot].synchronize()
                else:
                    self._slot_done[slot] = torch.cuda.Event()
                self.host_scores[slot].copy_(self.scores, non_blocking=True)
                self.host_image[slot].copy_(image, non_blocking=True)
                self._slot_done[slot].record()
                d2h += self.scores.numel() * 8 + image.numel() * 4
                scores, image = self.host_scores[slot], self.host_image[slot]
        self.launches, self.h2d_bytes, self.d2h_bytes = launches, h2d, d2h
        return scores, image, n_events

    def check_bounds(self):
        n = int(self.oob_total.item())
        if n != 0:
            raise IndexError("%d events are out of bounds for sensor_resolution %s" % (n, (self.H, self.W)))
