"""GPU parity, layer level: evk_conv2d_nhwc (the ConvLayer building block of every network; tcgen05 split-bf16
tensor-core kernel at precision 0, fp32 CUDA-core kernel at precision 1) against torch.nn.functional.conv2d in
fp32 on the CPU (the reference's own operator, model/submodules.py:14).

Tolerance: the split-bf16 scheme (hi*hi + lo*hi + hi*lo, fp32 accumulate in tensor memory) carries ~2^-16
relative operand error; bar = 3e-5 * max|ref| per layer (the 1e-4 end-to-end budget of north_star is checked in
test_gpu_networks.py); fp32 path: 5e-6 (summation order)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACT = {0: lambda t: t, 1: torch.relu, 2: torch.sigmoid, 3: torch.tanh}


def _conv(x_nchw, w, b, stride, pad, act, res, precision):
    from evreal_b200 import _lib
    lib = _lib.load()
    N, Cin, H, W = x_nchw.shape
    Cout, _, k, _ = w.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = x_nchw.permute(0, 2, 3, 1).contiguous().cuda()
    y = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32, device='cuda')
    r = res.permute(0, 2, 3, 1).contiguous().cuda() if res is not None else None
    wc, bc = w.contiguous(), (b.contiguous() if b is not None else None)
    _lib.check(lib.evk_conv2d_nhwc(_lib.ptr(x), N, H, W, Cin, ctypes.c_void_p(wc.data_ptr()),
                                   ctypes.c_void_p(bc.data_ptr()) if bc is not None else None, Cout, k, stride, pad, act,
                                   _lib.ptr(r) if r is not None else None, precision, _lib.ptr(y), _lib.stream_ptr()))
    return y.permute(0, 3, 1, 2).cpu()


CASES = [
    # N, Cin, H, W, Cout, k, stride, pad, act, residual
    (1, 64, 16, 24, 64, 3, 1, 1, 1, False),       # one exact 8x16 tile grid
    (2, 64, 23, 30, 128, 3, 1, 1, 0, True),       # ragged tiles + residual (ResidualBlock conv2 shape family)
    (1, 128, 46, 60, 64, 5, 1, 2, 1, False),      # decoder 5x5
    (1, 256, 23, 30, 256, 3, 1, 1, 1, False),     # resblock 256ch, two N tiles
    (1, 32, 40, 56, 64, 5, 2, 2, 1, False),       # encoder 0: Cin 32 -> 64B swizzle, stride 2
    (2, 64, 33, 47, 128, 5, 2, 2, 1, False),      # stride 2, odd sizes, batch 2
    (1, 64, 20, 20, 32, 5, 1, 2, 1, False),       # N tile 32 (decoder 2)
    (1, 64, 12, 20, 16, 3, 1, 1, 3, False),       # N tile 16, tanh
    (1, 64, 9, 11, 72, 3, 1, 1, 3, False),        # Cout 72 -> padded to 80 (HyperE2VID bases_net.3)
    (1, 1536, 10, 12, 128, 1, 1, 0, 1, False),    # 1x1 compositional conv (HyperE2VID)
    (1, 64, 4, 6, 64, 3, 1, 1, 1, False),         # tensor smaller than one tile
    (3, 96, 17, 19, 48, 3, 1, 1, 2, False),       # Cin 96 (32-chunks), Cout 48, sigmoid
]


@pytest.mark.parametrize('precision', [0, 1])
@pytest.mark.parametrize('case', CASES, ids=[str(c) for c in CASES])
def test_conv_layer_matches_torch(case, precision):
    N, Cin, H, W, Cout, k, stride, pad, act, use_res = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(x, w, b, stride, pad)
    res = torch.randn(ref.shape, generator=g) if use_res else None
    if res is not None:
        ref = ref + res
    ref = ACT[act](ref)
    got = _conv(x, w, b, stride, pad, act, res, precision)
    assert got.shape == ref.shape
    err = float((got - ref).abs().max())
    tol = (3e-5 if precision == 0 else 5e-6) * float(ref.abs().max())
    assert err <= tol, (err, tol)


def test_tensor_core_path_is_not_plain_bf16():
    """The split scheme must beat single-pass bf16 by orders of magnitude (guards against a silently dropped lo term)."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 128, 24, 32, generator=g)
    w = torch.randn(128, 128, 3, 3, generator=g) / (128 * 9) ** 0.5
    ref = F.conv2d(x, w, None, 1, 1)
    got = _conv(x, w, None, 1, 1, 0, None, 0)
    bf = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), None, 1, 1)
    err, err_bf16 = float((got - ref).abs().max()), float((bf - ref).abs().max())
    assert err < err_bf16 / 50, (err, err_bf16)


@pytest.mark.parametrize('case', [CASES[7], CASES[6], CASES[2], CASES[3], CASES[5]], ids=lambda c: str(c))
def test_tensor_core_conv_is_bit_reproducible(case):
    """The two MMA issuer warps must not introduce timing-dependent summation orders or pipeline races: 40 launches
    of the same layer give bit-identical outputs (a rare TMA completion reordering once showed up here as a 5e-5 blip)."""
    N, Cin, H, W, Cout, k, stride, pad, act, use_res = case
    g = torch.Generator().manual_seed(123)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    first = _conv(x, w, b, stride, pad, act, None, 0)
    ref = ACT[act](F.conv2d(x, w, b, stride, pad))
    assert float((first - ref).abs().max()) <= 3e-5 * float(ref.abs().max())
    for _ in range(40):
        again = _conv(x, w, b, stride, pad, act, None, 0)
        assert torch.equal(first, again)
