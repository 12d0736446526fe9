"""Host-side mirror of EVREAL's ``utils/eval_metrics.py`` on top of the CUDA metric kernels.

``BaseMetric`` / ``MseMetric`` / ``SsimMetric`` / ``EvalMetricsTracker`` keep the
reference's names, constructor arguments, gating rules and output files
(utils/eval_metrics.py:18-97, 162-350).  Images may be numpy arrays (drop-in use:
copied to the GPU) or CUDA tensors (resident fast path).  MSE and SSIM of one
frame come out of ONE fused kernel launch (``evk_mse_ssim``); the tracker shares
that launch between the two metric objects.
"""
import math
import traceback
from os.path import join

import numpy as np
import torch

from . import _lib
from .eval_utils import append_timestamp, append_result, ensure_dir, save_inferred_image, truncate_file


def _cuda_img(img):
    _lib.require_cuda()
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32))
    if not img.is_cuda:
        img = img.cuda(non_blocking=True)
    img = img.float()
    while img.dim() > 2 and img.shape[0] == 1:
        img = img[0]
    return img.contiguous()


def mse_ssim(img, ref, clip=False):
    """(mse, ssim) per image as a CUDA float64 tensor [n, 2]; img/ref: [H,W] or [n,H,W]."""
    a, b = _cuda_img(img), _cuda_img(ref)
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")     # skimage's message
    if a.dim() == 2:
        a, b = a[None], b[None]
    n, H, W = a.shape
    if H < 11 or W < 11:
        raise ValueError("win_size exceeds image extent.")                   # skimage's message (window 11)
    out = torch.empty((n, 2), dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().evk_mse_ssim(_lib.ptr(a), _lib.ptr(b), n, H, W, int(bool(clip)), _lib.ptr(out),
                                            _lib.stream_ptr(a.device)))
    return out


class BaseMetric:
    """Base class for quantitative evaluation metrics (utils/eval_metrics.py:18-74)."""

    def __init__(self, name, no_ref=False):
        self.scores = []
        self.name = name
        self.no_ref = no_ref
        self.updated = 0
        self.image_queue = []
        self.ref_queue = []
        self.batch_size = 4

    def reset(self):
        self.scores = []
        self.image_queue = []
        self.ref_queue = []
        self.updated = 0

    def finish_queue(self):
        self.updated = 0

    def get_num_updated(self):
        return self.updated

    def calculate(self, img, ref):
        raise NotImplementedError

    def update(self, img, ref=None):
        self.updated = 0
        score = self.calculate(img, ref)
        self.push(score)

    def push(self, score):
        if not isinstance(score, list):
            score = [score]
        for s in score:
            if math.isfinite(s) and not math.isnan(s):
                self.updated += 1
                self.scores.append(s)

    def get_num_scores(self):
        return len(self.scores)

    def get_all_scores(self):
        return self.scores

    def get_last_score(self):
        return self.scores[-1]

    def get_last_scores(self, n):
        return self.scores[-n:]

    def get_mean_score(self):
        if self.get_num_scores() == 0:
            return -1
        return sum(self.scores) / self.get_num_scores()

    def get_name(self):
        return self.name


class MseMetric(BaseMetric):
    """skimage.metrics.mean_squared_error(ref, img)  (utils/eval_metrics.py:77-84)."""
    column = 0

    def __init__(self):
        super().__init__(name='mse')

    def calculate(self, img, ref):
        return float(mse_ssim(img, ref)[0, self.column].item())


class SsimMetric(BaseMetric):
    """skimage.metrics.structural_similarity(ref, img, gaussian_weights=True, sigma=1.5,
    use_sample_covariance=False, data_range=1.0)  (utils/eval_metrics.py:87-97)."""
    column = 1

    def __init__(self, gaussian_weights=True, sigma=1.5, use_sample_covariance=False):
        super().__init__(name='ssim')
        if not gaussian_weights or sigma != 1.5 or use_sample_covariance:
            raise _lib.EvkError("only the reference's SSIM configuration (gaussian, sigma 1.5, population covariance) is built")
        self.gaussian_weights = gaussian_weights
        self.sigma = sigma
        self.use_sample_covariance = use_sample_covariance

    def calculate(self, img, ref):
        return float(mse_ssim(img, ref)[0, self.column].item())


def create_metric(name):
    if name == 'mse':
        return MseMetric()
    if name == 'ssim':
        return SsimMetric()
    if name in ('lpips', 'lpips-vgg'):
        from .lpips import LpipsMetric
        return LpipsMetric(name)
    return None


class EvalMetricsTracker:
    """utils/eval_metrics.py:162-350: clip, (optional PNG), gate by evaluation window and timestamp
    tolerance, update metrics, append per-frame scores to text files.

    ``defer=True`` keeps every score on the GPU until ``finalize`` (one device->host copy per
    sequence instead of one per frame); files and score lists are identical afterwards.
    Histogram equalisation (utils/eval_metrics.py:326-350): 'global' runs on the GPU (evk_equalize_hist, skimage's
    published algorithm -- parity unpinned, scikit-image is not installable offline), 'clahe' is the reference's own
    cv2.createCLAHE call on the host, 'local' (skimage rank filter over a disk of radius 55) runs on the GPU
    (evk_equalize_local; parity unpinned like 'global').
    """

    def __init__(self, save_images=False, save_processed_images=False, output_dir=None, hist_eq='none',
                 quan_eval_metric_names=None, quan_eval_start_time=0, quan_eval_end_time=float('inf'),
                 quan_eval_ts_tol_ms=float('inf'), has_reference_frames=False, color=False, defer=False,
                 write_files=True):
        if quan_eval_metric_names is None:
            quan_eval_metric_names = ['mse', 'ssim', 'lpips']
        if hist_eq not in ('none', 'global', 'clahe', 'local'):
            raise ValueError(f"Unrecognized histogram equalization argument: {hist_eq}")
        self.save_images = save_images
        self.save_processed_images = save_processed_images
        if hist_eq == 'none' and self.save_processed_images:
            print("Can not save processed images when hist_eq is none")
            self.save_processed_images = False
        self.output_dir = output_dir
        self.hist_eq = hist_eq
        self.quan_eval_start_time = quan_eval_start_time
        self.quan_eval_end_time = quan_eval_end_time
        self.quan_eval_ts_tol_ms = quan_eval_ts_tol_ms
        self.has_reference_frames = has_reference_frames
        self.color = color
        self.defer = defer
        self.write_files = write_files and output_dir is not None
        self.quan_eval_indices = []
        self._pending = []          # deferred mode: CUDA [1,2] score tensors, one per evaluated frame

        self.metrics = []
        for metric_name in quan_eval_metric_names:
            m = create_metric(metric_name)
            if m is None:
                print("Unknown metric " + metric_name)
            else:
                self.metrics.append(m)
        if not self.has_reference_frames:
            self.metrics = [m for m in self.metrics if m.no_ref]
        self.only_no_ref = all([m.no_ref for m in self.metrics])
        self.reset()

    def reset(self):
        if self.write_files:
            self.setup_output_folders_and_files()
        for metric in self.metrics:
            metric.reset()
        self._pending = []

    def save_new_scores(self, idx, metric):
        if not self.write_files:
            return
        num_updated = metric.get_num_updated()
        if num_updated > 0:
            last_scores = metric.get_last_scores(num_updated)
            indices = self.quan_eval_indices[-num_updated:]
            append_result(self.get_metric_file_path(metric), indices, last_scores)

    def _flush_deferred(self):
        if not self._pending:
            return
        host = torch.cat(self._pending).cpu().numpy()      # one D2H for the whole sequence
        self._pending = []
        indices = self.quan_eval_indices[-len(host):]
        for metric in self.metrics:
            col = getattr(metric, 'column', None)
            if col is None:
                continue
            metric.updated = 0
            vals = [float(v) for v in host[:, col]]
            metric.push(vals)
            if self.write_files and metric.get_num_updated() > 0:
                # BaseMetric.push drops non-finite scores: the index column drops the same rows
                kept = [i for i, v in zip(indices, vals) if math.isfinite(v)]
                append_result(self.get_metric_file_path(metric), kept, metric.get_last_scores(metric.get_num_updated()))

    def finalize(self, idx):
        self._flush_deferred()
        for metric in self.metrics:
            if getattr(metric, 'column', None) is not None and self.defer:
                continue
            metric.finish_queue()
            self.save_new_scores(idx, metric)

    def update_quantitative_metrics(self, idx, img, ref):
        self.quan_eval_indices.append(idx)
        fused = None
        for metric in self.metrics:
            try:
                col = getattr(metric, 'column', None)
                if col is not None and self.has_reference_frames:
                    if fused is None:
                        fused = mse_ssim(img, ref)           # shared by MseMetric and SsimMetric
                    if self.defer:
                        continue
                    metric.updated = 0
                    metric.push(float(fused[0, col].item()))
                elif not self.has_reference_frames or metric.no_ref:
                    metric.update(img)
                else:
                    metric.update(img, ref)
                self.save_new_scores(idx, metric)
            except Exception as e:
                print("Exception in metric " + metric.get_name() + ": " + str(e))
                print(traceback.format_exc())
                metric.reset()
        if self.defer and fused is not None:
            self._pending.append(fused)

    def update(self, idx, img, ref, img_ts, ref_ts=None):
        if ref_ts is None:
            ref_ts = img_ts
        if self.write_files:
            append_timestamp(self.get_timestamps_file_path(), idx, img_ts)

        # clip images (utils/eval_metrics.py:253-255)
        if isinstance(img, np.ndarray):
            img = np.clip(img, 0.0, 1.0)
        else:
            img = img.clamp(0.0, 1.0)
        if self.has_reference_frames:
            ref = np.clip(ref, 0.0, 1.0) if isinstance(ref, np.ndarray) else ref.clamp(0.0, 1.0)

        if self.save_images and self.output_dir is not None:
            save_inferred_image(self.output_dir, img if isinstance(img, np.ndarray) or img.dim() != 2 else img, idx)

        img = self.histogram_equalization(img)
        if self.has_reference_frames:
            ref = self.histogram_equalization(ref)
        if self.save_processed_images and self.output_dir is not None:
            save_inferred_image(self.processed_output_dir, img, idx)

        inside_eval_cut = self.quan_eval_start_time <= img_ts <= self.quan_eval_end_time
        img_ref_time_diff_ms = abs(ref_ts - img_ts) * 1000
        inside_eval_ts_tolerance = img_ref_time_diff_ms <= self.quan_eval_ts_tol_ms
        if self.only_no_ref:
            inside_eval_ts_tolerance = True
        if inside_eval_cut and inside_eval_ts_tolerance and not self.color:
            self.update_quantitative_metrics(idx, img, ref)

    def save_custom_metric(self, idx, metric_name, metric_value, is_int=False):
        if not self.write_files:
            return
        metric_file_path = join(self.output_dir, metric_name + '.txt')
        if idx == 0:
            truncate_file(metric_file_path)
        append_result(metric_file_path, idx, metric_value, is_int)

    def get_num_quan_evaluations(self):
        return len(self.quan_eval_indices)

    def get_mean_scores(self):
        return {metric.get_name(): metric.get_mean_score() for metric in self.metrics}

    def get_timestamps_file_path(self):
        return join(self.output_dir, 'timestamps.txt')

    def get_metric_file_path(self, metric):
        return join(self.output_dir, metric.get_name() + '.txt')

    def setup_output_folders_and_files(self):
        ensure_dir(self.output_dir)
        if self.save_processed_images:
            self.processed_output_dir = self.output_dir + "_processed"
            ensure_dir(self.processed_output_dir)
        truncate_file(self.get_timestamps_file_path())
        for metric in self.metrics:
            truncate_file(self.get_metric_file_path(metric))

    def histogram_equalization(self, img):
        """utils/eval_metrics.py:326-350 on a clipped [H, W] frame (numpy or CUDA tensor; same type out)."""
        if self.hist_eq == 'none':
            return img
        if self.hist_eq == 'global':
            was_numpy = isinstance(img, np.ndarray)
            x = _cuda_img(img)
            out = torch.empty_like(x)
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().evk_equalize_hist(_lib.ptr(x), _lib.ptr(out), 1, x.numel(), 0, _lib.stream_ptr(x.device)))
            return out.cpu().numpy() if was_numpy else out
        if self.hist_eq == 'local':
            was_numpy = isinstance(img, np.ndarray)
            x = _cuda_img(img)
            out = torch.empty_like(x)
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().evk_equalize_local(_lib.ptr(x), _lib.ptr(out), 1, int(x.shape[-2]), int(x.shape[-1]), 55, 0,
                                                          _lib.stream_ptr(x.device)))
            return out.cpu().numpy() if was_numpy else out
        if self.hist_eq == 'clahe':
            import cv2
            was_numpy = isinstance(img, np.ndarray)
            a = img if was_numpy else img.cpu().numpy()
            # img_as_ubyte of a float image in [0, 1]: round(v * 255); img_as_float32 of uint8: v / 255
            a8 = np.round(np.clip(a, 0.0, 1.0) * 255).astype(np.uint8)
            a8 = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(a8)
            out = a8.astype(np.float32) / np.float32(255)
            return out if was_numpy else torch.from_numpy(out).to(img.device)
        raise ValueError(f"Unrecognized histogram equalization argument: {self.hist_eq}")

    def create_video(self):
        print("create_video is outside the accelerated hot path (create_video:false in all shipped eval configs)")

    def create_processed_video(self):
        print("create_processed_video is outside the accelerated hot path")
