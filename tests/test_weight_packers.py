"""CPU: the host-side weight re-packers inside libevreal_b200.so (poly.cu / conv_tc.cu, reached through the host-only C-ABI
entry evk_pack_layer_weights -- no CUDA call is made) against torch's own operators in float64: a convolution evaluated with the
packed matrix in the kernel's GEMM view must equal the reference layer it replaces."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from evreal_b200 import _lib


def _pack(kind, w, group=1, cap=1 << 22):
    lib = _lib.load()
    w32 = np.ascontiguousarray(w.numpy().astype(np.float32))
    out = np.zeros(cap, dtype=np.float32)
    n = ctypes.c_int64(0)
    Cout, Cin, kh, kw = w32.shape
    _lib.check(lib.evk_pack_layer_weights(kind, w32.ctypes.data_as(ctypes.c_void_p), Cout, Cin, kh, kw, group,
                                          out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n)))
    return torch.from_numpy(out[:n.value].copy()).double()


def _w(Cout, Cin, kh, kw, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(Cout, Cin, kh, kw, generator=g) * 0.5).float().double()      # float32-representable values


def test_phase_stacked_weights_reproduce_upsample_conv_in_the_interior():
    Cin, Cout, H, W = 8, 4, 7, 9
    w = _w(Cout, Cin, 5, 5, 0)
    x = torch.randn(1, Cin, H, W, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False), w, padding=2)
    m = _pack(0, w).reshape(5, 5, Cin, 4, Cout)                      # [ty][tx][c][phase][n]
    xp = F.pad(x, (2, 2, 2, 2), mode='replicate')
    out = torch.zeros_like(ref)
    for ph in range(4):
        wc = m[:, :, :, ph, :].permute(3, 2, 0, 1).contiguous()      # [n][c][ty][tx]
        out[:, :, ph >> 1::2, ph & 1::2] = F.conv2d(xp, wc)
    assert float((out - ref)[:, :, 2:-2, 2:-2].abs().max()) < 1e-5   # composite weights are stored in float32
    # the tap rows / columns the kernel skips per tile hold exact zeros
    for ph in range(4):
        a, b = ph >> 1, ph & 1
        assert float(m[4 if a == 0 else 0, :, :, ph, :].abs().max()) == 0.0
        assert float(m[:, 4 if b == 0 else 0, :, ph, :].abs().max()) == 0.0


def test_border_line_weights_are_the_negated_sums_of_the_outside_taps():
    Cin, Cout = 8, 4
    w = _w(Cout, Cin, 5, 5, 1)
    m = _pack(1, w).reshape(2, 5, Cin, 4, Cout)                      # [v][s][c][line][n]
    dsets = [(-2, -1), (-2,), (2,), (1, 2)]
    for l in range(4):
        for s in range(5):
            h = -sum(w[:, :, d + 2, s] for d in dsets[l])            # horizontal border: rows outside, per column offset
            v = -sum(w[:, :, s, d + 2] for d in dsets[l])            # vertical border: columns outside, per row offset
            assert torch.allclose(m[0, s, :, l, :], h.t(), atol=1e-6, rtol=0)
            assert torch.allclose(m[1, s, :, l, :], v.t(), atol=1e-6, rtol=0)


def test_pixel_pair_weights_reproduce_the_stride2_layer():
    C, Co, H, W = 4, 8, 9, 12
    w = _w(Co, C, 5, 5, 2)
    x = torch.randn(2, C, H, W, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=2, padding=2)
    m = _pack(2, w).reshape(5, 3, 2 * C, Co)                         # [r][pair][slot*C + c][n]
    xpair = x.permute(0, 2, 3, 1).reshape(2, H, W // 2, 2 * C).permute(0, 3, 1, 2)
    got = F.conv2d(xpair, m.permute(3, 2, 0, 1).contiguous(), stride=(2, 1), padding=(2, 1))
    assert float((got - ref).abs().max()) < 1e-12


def test_window_weights_reproduce_the_3x3_layer_with_two_pixels_per_row():
    C, T, Co, H, W, G = 16, 2, 4, 5, 8, 2                            # cat(x, h): two 16-channel tensors
    w = _w(Co, T * C, 3, 3, 3)
    x = torch.randn(1, T * C, H, W, dtype=torch.float64)
    ref = F.conv2d(x, w, padding=1)
    m = _pack(3, w, group=G).reshape(3, T, 4, C, G, Co)              # [r][tensor][slot][c][g][n]
    xp = F.pad(x, (1, 1, 1, 1))                                      # pixel x at column x + 1; rows: TMA out-of-bounds zero fill
    out = torch.zeros_like(ref)
    for j in range(W // G):
        for y in range(H):
            win = xp[0, :, y:y + 3, G * j:G * j + 4].reshape(T, C, 3, 4)      # [tensor][c][r][slot]
            out[0, :, y, G * j:G * j + G] = torch.einsum('tcrs,rtscgn->ng', win, m)
    assert float((out - ref).abs().max()) < 1e-12
    assert float(m[:, :, 3, :, 0, :].abs().max()) == 0.0 and float(m[:, :, 0, :, 1, :].abs().max()) == 0.0   # unused window slots


def test_bad_arguments_fail_with_a_message():
    lib = _lib.load()
    n = ctypes.c_int64(0)
    buf = np.zeros(16, dtype=np.float32)
    w = np.zeros((4, 8, 5, 5), dtype=np.float32)
    rc = lib.evk_pack_layer_weights(0, w.ctypes.data_as(ctypes.c_void_p), 4, 8, 5, 5, 1, buf.ctypes.data_as(ctypes.c_void_p), 16, ctypes.byref(n))
    assert rc != 0 and n.value == 25 * 8 * 4 * 4 and b'needs' in lib.evk_last_error()
    rc = lib.evk_pack_layer_weights(9, w.ctypes.data_as(ctypes.c_void_p), 4, 8, 5, 5, 1, buf.ctypes.data_as(ctypes.c_void_p), 16, ctypes.byref(n))
    assert rc != 0 and b'unknown kind' in lib.evk_last_error()


def test_mixed_operand_weights_decode_to_the_layer():
    """The mixed-operand form of a plain layer (fp16 product + two fp8 products; conv.cuh, ConvParams::mixed): the fp16 part plus
    the e4m3 remainder reproduce every weight to 2^-14 of its output channel's largest weight, the e4m3 copy to 2^-4 of the
    value (4 significand bits) down to the format's subnormal range, with channel magnitudes spread over four decades."""
    Cout, Cin, k = 48, 64, 3
    w = _w(Cout, Cin, k, k, 5) * torch.logspace(-3, 1, Cout, dtype=torch.float64).view(-1, 1, 1, 1)
    w = w.float().double()
    K = k * k * Cin
    d = _pack(4, w, cap=3 * K * Cout).reshape(3, k, k, Cin, Cout).permute(0, 4, 3, 1, 2)        # [part][n][c][r][q]
    cmax = w.abs().amax(dim=(1, 2, 3), keepdim=True)
    assert float(((d[0] + d[1] - w).abs() / cmax).max()) <= 2.0 ** -14
    big = w.abs() >= cmax * 2.0 ** -11                                # e4m3 normal range under the channel's scale
    assert float(((d[2] - w).abs() / w.abs())[big].max()) <= 2.0 ** -4 + 1e-9
    assert float(((d[2] - w).abs() / cmax)[~big].max()) <= 2.0 ** -14            # below it: the subnormal step
