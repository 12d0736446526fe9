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


def _append(path, rows, fmt):
    with open(path, 'a', encoding="utf-8") as f:
        f.writelines(fmt.format(k, v) for k, v in rows)


def append_timestamp(path, description, timestamp):
    """'<index> <timestamp with 15 decimals>' (utils/eval_utils.py:57-59)."""
    _append(path, [(description, timestamp)], '{} {:.15f}\n')


def append_result(path, description, result, is_int=False):
    """'<index> <score>' lines, 5 decimals unless is_int; a list of results pairs up with a list of indices
    (utils/eval_utils.py:62-77)."""
    rows = list(zip(description, result)) if isinstance(result, list) else [(description, result)]
    _append(path, rows, '{} {}\n' if is_int else '{} {:.5f}\n')


def save_inferred_image(folder, image, idx):
    """PNG writer (utils/eval_utils.py:80-84); cv2 is imported lazily -- it is not on the hot path."""
    import cv2
    png_path = join(folder, 'frame_{:010d}.png'.format(idx))
    cv2.imwrite(png_path, np.round(np.asarray(image) * 255).astype(np.uint8))
