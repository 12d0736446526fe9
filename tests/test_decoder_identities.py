"""CPU: the algebraic identities the re-shaped tensor-core layers rest on, checked against torch's own operators
(float64, so the identities are exact up to rounding).  The CUDA packers in evreal_b200/csrc (pack_weights_phase4,
pack_weights_ring, pack_weights_pixel_pair, pack_weights_window) implement exactly these index maps; the GPU parity
tests (tests/test_gpu_networks.py) then pin the kernels against the reference's frames.

 * phase-stacked UpsampleConvLayer (model/submodules.py:69-97): conv5x5_pad2(bilinear_x2(x)) == four 5x5 phase
   convolutions of the replicate-padded low-resolution map minus a border correction that is a 1x5 convolution along
   each border of one line of the extended upsampled map;
 * first encoder (stride 2, 5x5) == a 5x3 stride-(2,1) convolution over pixel pairs ([N,H,W,C] viewed as [N,H,W/2,2C]);
 * 3x3 stride-1 convolution with two output pixels per row over 4-pixel windows of a row-padded tensor (FireNet).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _phase_coef(a):
    """c[d+2][t]: weight of low-resolution sample (i + t - 2) in upsampled sample (2i + a + d)  (poly.cu phase_coef)."""
    c = np.zeros((5, 5))
    for d in range(-2, 3):
        s = a + d
        m = (s + 4) // 2 - 2
        r = (s + 4) & 1
        if r == 0:
            c[d + 2][m - 1 + 2] += 0.25
            c[d + 2][m + 2] += 0.75
        else:
            c[d + 2][m + 2] += 0.75
            c[d + 2][m + 1 + 2] += 0.25
    return torch.tensor(c)


def _u_ext(xp, Y, X):
    """extended upsampled map on the replicate-padded input xp (pad 2), valid for Y in [-2, 2H+2)."""
    my, ry, mx, rx = (Y + 4) // 2 - 2, (Y + 4) & 1, (X + 4) // 2 - 2, (X + 4) & 1
    rowA, colA = (my if ry else my - 1) + 2, (mx if rx else mx - 1) + 2
    wy, wx = (0.75 if ry else 0.25), (0.75 if rx else 0.25)
    v = 0
    for dy in range(2):
        for dx in range(2):
            v = v + (1 - wy if dy else wy) * (1 - wx if dx else wx) * xp[0, :, rowA + dy, colA + dx]
    return v


def test_phase_stacked_upsample_conv_with_border_lines():
    torch.manual_seed(0)
    Cin, Cout, H, W = 4, 3, 6, 7
    x = torch.randn(1, Cin, H, W, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 5, 5, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False), w, padding=2)
    xp = F.pad(x, (2, 2, 2, 2), mode='replicate')
    out = torch.zeros_like(ref)
    for a in range(2):
        for b in range(2):
            wc = torch.einsum('oiyx,yt,xs->oits', w, _phase_coef(a), _phase_coef(b))
            out[:, :, a::2, b::2] = F.conv2d(xp, wc)
            # the tap rows / columns the kernel skips per tile are identically zero
            assert float(wc[:, :, 4 if a == 0 else 0, :].abs().max()) == 0.0
            assert float(wc[:, :, :, 4 if b == 0 else 0].abs().max()) == 0.0
    assert float((out - ref)[:, :, 2:-2, 2:-2].abs().max()) < 1e-12          # interior: no correction needed
    Ho, Wo = 2 * H, 2 * W
    dsets = [(-2, -1), (-2,), (2,), (1, 2)]
    corr = torch.zeros_like(out)
    for l in range(4):                                                       # horizontal border lines
        Yo, src = (l if l < 2 else Ho - 4 + l), (-1 if l < 2 else Ho)
        for X in range(Wo):
            for s in range(5):
                corr[0, :, Yo, X] += sum(w[:, :, d + 2, s] for d in dsets[l]) @ _u_ext(xp, src, X + s - 2)
    for l in range(4):                                                       # vertical: rows outside belong to the horizontal pass
        Xo, src = (l if l < 2 else Wo - 4 + l), (-1 if l < 2 else Wo)
        for Y in range(Ho):
            for s in range(5):
                if 0 <= Y + s - 2 < Ho:
                    corr[0, :, Y, Xo] += sum(w[:, :, s, d + 2] for d in dsets[l]) @ _u_ext(xp, Y + s - 2, src)
    assert float((out - corr - ref).abs().max()) < 1e-12


def test_stride2_5x5_as_5x3_over_pixel_pairs():
    torch.manual_seed(1)
    C, Co, H, W = 6, 5, 9, 12
    x = torch.randn(2, C, H, W, dtype=torch.float64)
    w = torch.randn(Co, C, 5, 5, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=2, padding=2)
    # [N,H,W,C] -> [N,H,W/2,2C]: channel index = slot * C + c
    xpair = x.permute(0, 2, 3, 1).reshape(2, H, W // 2, 2 * C).permute(0, 3, 1, 2)
    w2 = torch.zeros(Co, 2 * C, 5, 3, dtype=torch.float64)
    for q in range(5):                                                       # tap q = pair q // 2, slot q % 2
        w2[:, (q % 2) * C:(q % 2 + 1) * C, :, q // 2] = w[:, :, :, q]
    got = F.conv2d(xpair, w2, stride=(2, 1), padding=(2, 1))
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 1e-12


def test_3x3_two_pixels_per_row_over_padded_windows():
    torch.manual_seed(2)
    C, Co, H, W, G, P = 4, 3, 5, 8, 2, 4                                    # P pixels per window, G outputs per window
    x = torch.randn(1, C, H, W, dtype=torch.float64)
    w = torch.randn(Co, C, 3, 3, dtype=torch.float64)
    ref = F.conv2d(x, w, padding=1)
    xp = F.pad(x, (1, 1, 0, 0))                                              # row-padded: pixel x at column x + 1
    out = torch.zeros_like(ref)
    for j in range(W // G):
        win = xp[0, :, :, G * j:G * j + P]                                   # [C, H, P]: the window of GEMM row j (all image rows)
        win = F.pad(win, (0, 0, 1, 1))                                       # vertical zero padding (TMA out-of-bounds fill)
        for g in range(G):                                                   # output pixel g reads tap q from slot g + q
            for y in range(H):
                acc = torch.zeros(Co, dtype=torch.float64)
                for r in range(3):
                    for q in range(3):
                        acc += w[:, :, r, q] @ win[:, y + r, g + q]
                out[0, :, y, G * j + g] = acc
    assert float((out - ref).abs().max()) < 1e-12
