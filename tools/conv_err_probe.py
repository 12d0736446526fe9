"""Error of one tensor-core convolution case against torch fp32 (CPU) under different forced tilings."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch, torch.nn.functional as F
from test_gpu_conv import _conv, ACT
case = (1, 64, 12, 20, 16, 3, 1, 1, 3, False)
N, Cin, H, W, Cout, k, stride, pad, act, use_res = case
g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
x = torch.randn(N, Cin, H, W, generator=g)
w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
b = torch.randn(Cout, generator=g) * 0.1
pre = F.conv2d(x, w, b, stride, pad)
pre64 = F.conv2d(x.double(), w.double(), b.double(), stride, pad)
print('torch fp32 vs fp64 pre-activation max err', float((pre.double() - pre64).abs().max()))
for act_id in (0, 3):
    ref = ACT[act_id](pre)
    for env in ({}, {'EVK_TC_CS': '1'}, {'EVK_TC_ISSUERS': '1'}, {'EVK_TC_UX': '0'}, {'EVK_TC_UX': '1'}, {'EVK_TC_OWN_ACC': '0'}):
        for kk in ('EVK_TC_CS', 'EVK_TC_ISSUERS', 'EVK_TC_UX', 'EVK_TC_OWN_ACC'):
            os.environ.pop(kk, None)
        os.environ.update(env)
        got = _conv(x, w, b, stride, pad, act_id, None, 0)
        err = (got - ref).abs()
        print('act', act_id, env, 'max err %.3e' % float(err.max()), 'rms %.3e' % float(err.pow(2).mean().sqrt()),
              'vs fp64: %.3e' % float((got.double() - ACT[act_id](pre64)).abs().max()))
got1 = _conv(x, w, b, stride, pad, 0, None, 1)
print('fp32 simt max err', float((got1 - pre).abs().max()))
