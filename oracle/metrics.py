"""Oracle: per-frame quality metrics (stage 3) and post-processing normalisation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, in /root/reference:
  utils/eval_metrics.py:77-84     MseMetric   -> skimage.metrics.mean_squared_error(ref, img)
  utils/eval_metrics.py:87-97     SsimMetric  -> skimage.metrics.structural_similarity(ref, img,
                                    gaussian_weights=True, sigma=1.5, use_sample_covariance=False, data_range=1.0)
  utils/eval_metrics.py:100-156   pyiqa LPIPS via queue of 4
  utils/eval_metrics.py:244-273   clip + gating
  utils/eval_utils.py:15-35       robust percentile normalisation
  eval.py:380-395                 post_process_normalization

Third-party arithmetic (absent from /root/reference, unpinned in requirements.txt:3,7):
  * scikit-image >= 0.19 ``structural_similarity``: published algorithm restated
    (SURVEY A.3).  scikit-image filters with ``scipy.ndimage.gaussian_filter``;
    ``gaussian_blur_f32`` below restates scipy's separable correlate1d
    (float64 accumulation per 1-D pass, float32 storage between passes, axis 0
    first) and is pinned against scipy itself in tests/test_oracle_metrics.py.
    PARITY UNPINNED against scikit-image proper (package not available offline).
  * pyiqa LPIPS: PARITY UNPINNED (package and weights not available offline);
    ``lpips_oracle`` restates the public richzhang/pyiqa formula for a caller
    supplied weight dict.
"""
import numpy as np
import torch
import torch.nn.functional as F

SSIM_SIGMA = 1.5
SSIM_TRUNCATE = 3.5
SSIM_RADIUS = int(SSIM_TRUNCATE * SSIM_SIGMA + 0.5)      # 5 -> 11 taps
K1, K2 = 0.01, 0.03


def gaussian_taps(sigma=SSIM_SIGMA, radius=SSIM_RADIUS):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) (float64)."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _correlate1d_sym(a, taps, axis):
    """scipy's symmetric-kernel branch of correlate1d, mode='reflect':
    tmp = x[l]*w[r] ; tmp += (x[l-j] + x[l+j]) * w[r-j], all in float64;
    result rounded to the array dtype (float32)."""
    r = len(taps) // 2
    a64 = np.moveaxis(a, axis, -1).astype(np.float64)
    pad = np.pad(a64, [(0, 0)] * (a64.ndim - 1) + [(r, r)], mode='symmetric')
    n = a64.shape[-1]
    out = pad[..., r:r + n] * taps[r]
    for j in range(r, 0, -1):           # scipy walks ll = -size1 .. -1
        out = out + (pad[..., r - j:r - j + n] + pad[..., r + j:r + j + n]) * taps[r - j]
    return np.moveaxis(out.astype(a.dtype), -1, axis)


def gaussian_blur_f32(a):
    a = np.asarray(a, dtype=np.float32)
    taps = gaussian_taps()
    return _correlate1d_sym(_correlate1d_sym(a, taps, 0), taps, 1)


def mse_oracle(img, ref):
    """skimage.metrics.mean_squared_error: mean((a-b)**2, dtype=float64) of float32 inputs."""
    img = np.asarray(img, dtype=np.float32)
    ref = np.asarray(ref, dtype=np.float32)
    return float(np.mean((ref - img) ** 2, dtype=np.float64))


def ssim_oracle(img, ref, data_range=1.0):
    """skimage structural_similarity with the reference's arguments."""
    x = np.asarray(ref, dtype=np.float32)        # reference passes (ref, img)
    y = np.asarray(img, dtype=np.float32)
    ux, uy = gaussian_blur_f32(x), gaussian_blur_f32(y)
    uxx, uyy, uxy = gaussian_blur_f32(x * x), gaussian_blur_f32(y * y), gaussian_blur_f32(x * y)
    vx = uxx - ux * ux                            # cov_norm = 1.0
    vy = uyy - uy * uy
    vxy = uxy - ux * uy
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    A1, A2 = 2 * ux * uy + C1, 2 * vxy + C2
    B1, B2 = ux ** 2 + uy ** 2 + C1, vx + vy + C2
    S = (A1 * A2) / (B1 * B2)
    p = SSIM_RADIUS                               # (win_size - 1) // 2
    return float(S[p:-p, p:-p].mean(dtype=np.float64))


def robust_normalize_oracle(img, q_min=1, q_max=99):
    """utils/eval_utils.py:15-35 (np.percentile, linear interpolation, float32 result)."""
    img = np.asarray(img, dtype=np.float32)
    lo = np.percentile(img.ravel(), q_min)
    hi = np.percentile(img.ravel(), q_max)
    return (img - lo) / (hi - lo)


def post_process_oracle(img, norm):
    """eval.py:380-395."""
    if norm == 'robust':
        return robust_normalize_oracle(img, 1, 99)
    if norm == 'standard':
        return robust_normalize_oracle(img, 0, 100)
    if norm == 'exprobust':
        return robust_normalize_oracle(np.exp(img), 1, 99)
    if norm == 'none':
        return img
    raise ValueError("Unrecognized normalization argument: %s" % norm)


def equalize_hist_oracle(img, nbins=256):
    """EvalMetricsTracker.histogram_equalization, hist_eq == 'global' (utils/eval_metrics.py:326-331):
    skimage.exposure.equalize_hist -> img_as_float32.  PARITY UNPINNED (scikit-image is not installable offline); its
    published algorithm: hist, edges = np.histogram(img, nbins) over [min, max]; cdf = cumsum(hist) / numel;
    out = np.interp(img, bin centres, cdf)."""
    img = np.asarray(img, dtype=np.float32)
    hist, edges = np.histogram(img.ravel(), bins=nbins)
    centres = (edges[:-1] + edges[1:]) / 2.0
    cdf = hist.cumsum()
    cdf = cdf / float(cdf[-1])
    return np.interp(img.ravel(), centres, cdf).reshape(img.shape).astype(np.float32)


def equalize_local_oracle(img, radius=55):
    """EvalMetricsTracker.histogram_equalization, hist_eq == 'local' (utils/eval_metrics.py:332-339):
    img_as_float32(skimage.filters.rank.equalize(img_as_ubyte(img), footprint=disk(radius))).  PARITY UNPINNED (scikit-image is
    not installable offline); its published algorithm (filters/rank/generic_cy.pyx::_kernel_equalize): for every pixel, over the
    footprint pixels inside the image (pop of them), uint8(255 * #{v <= g} / pop), g = the pixel's own grey level; disk(r) =
    {dx^2 + dy^2 <= r^2}; img_as_ubyte rounds v * 255 half-to-even in float32, img_as_float32 multiplies by float32(1 / 255)."""
    img = np.asarray(img, dtype=np.float32)
    u8 = np.clip(np.rint(img * np.float32(255)), 0, 255).astype(np.int32)
    H, W = u8.shape
    pad = np.full((H + 2 * radius, W + 2 * radius), -1, dtype=np.int32)
    pad[radius:radius + H, radius:radius + W] = u8
    pop = np.zeros((H, W), dtype=np.int64)
    le = np.zeros((H, W), dtype=np.int64)
    for dy in range(-radius, radius + 1):
        w = int(np.floor(np.sqrt(radius * radius - dy * dy)))
        for dx in range(-w, w + 1):
            q = pad[radius + dy:radius + dy + H, radius + dx:radius + dx + W]
            valid = q >= 0
            pop += valid
            le += valid & (q <= u8)
    out = np.where(pop > 0, (255 * le) / np.maximum(pop, 1).astype(np.float64), 0.0).astype(np.uint8)      # C cast: truncation
    return out.astype(np.float32) * np.float32(1.0 / 255.0)


def quantize_u8_oracle(img):
    """save_inferred_image (utils/eval_utils.py:80-84) after the tracker's clip: uint8(np.round(clip(img, 0, 1) * 255))."""
    return np.round(np.clip(np.asarray(img, dtype=np.float32), 0.0, 1.0) * 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# LPIPS (PARITY UNPINNED -- see module docstring)
# ----------------------------------------------------------------------------
LPIPS_SHIFT = (-.030, -.088, -.188)
LPIPS_SCALE = (.458, .448, .450)
# (out_channels, kernel, stride, pad, maxpool_before) per conv; taps after ReLU of each listed conv
ALEX_CFG = [(64, 11, 4, 2, False), (192, 5, 1, 2, True), (384, 3, 1, 1, True), (256, 3, 1, 1, False),
            (256, 3, 1, 1, False)]
ALEX_TAPS = (0, 1, 2, 3, 4)
VGG_CFG = [(64, 3, 1, 1, False), (64, 3, 1, 1, False),
           (128, 3, 1, 1, True), (128, 3, 1, 1, False),
           (256, 3, 1, 1, True), (256, 3, 1, 1, False), (256, 3, 1, 1, False),
           (512, 3, 1, 1, True), (512, 3, 1, 1, False), (512, 3, 1, 1, False),
           (512, 3, 1, 1, True), (512, 3, 1, 1, False), (512, 3, 1, 1, False)]
VGG_TAPS = (1, 3, 6, 9, 12)


def lpips_backbone_cfg(net):
    return (ALEX_CFG, ALEX_TAPS, 3) if net == 'alex' else (VGG_CFG, VGG_TAPS, 2)


def random_lpips_weights(net='alex', seed=0):
    """Seeded stand-in for the unavailable pretrained backbone + linear heads."""
    cfg, taps, _ = lpips_backbone_cfg(net)
    g = torch.Generator().manual_seed(seed)
    w = {}
    cin = 3
    for i, (cout, k, _, _, _) in enumerate(cfg):
        w['conv%d.weight' % i] = torch.randn(cout, cin, k, k, generator=g) * (1.6 / (cin * k * k) ** 0.5)
        w['conv%d.bias' % i] = torch.randn(cout, generator=g) * 0.05
        cin = cout
    for j, t in enumerate(taps):
        w['lin%d.weight' % j] = torch.rand(1, cfg[t][0], 1, 1, generator=g) / cfg[t][0]   # non-negative
    return w


@torch.no_grad()
def lpips_oracle(img, ref, w, net='alex'):
    """img, ref: [N,3,H,W] in [0,1].  Returns [N] scores (public LPIPS v0.1 formula)."""
    cfg, taps, pool_k = lpips_backbone_cfg(net)
    shift = torch.tensor(LPIPS_SHIFT).view(1, 3, 1, 1)
    scale = torch.tensor(LPIPS_SCALE).view(1, 3, 1, 1)

    def feats(x):
        x = (2 * x - 1 - shift) / scale
        out = []
        for i, (_, k, s, p, pool) in enumerate(cfg):
            if pool:
                x = F.max_pool2d(x, pool_k, 2)
            x = torch.relu(F.conv2d(x, w['conv%d.weight' % i], w['conv%d.bias' % i], s, p))
            if i in taps:
                out.append(x)
        return out

    total = 0
    for j, (a, b) in enumerate(zip(feats(img), feats(ref))):
        a = a / (a.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
        b = b / (b.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
        d = (a - b) ** 2
        total = total + F.conv2d(d, w['lin%d.weight' % j]).mean((2, 3))
    return total.view(-1)
