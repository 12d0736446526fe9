"""Voxelizer probe: cfg-5 points, CUDA-event timing per call (run under ncu for per-kernel times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evreal_b200 import _lib
lib = _lib.load()
Hv, Wv, bins = 480, 640, 5
n = int(os.environ.get('N', '4000000'))
g = torch.Generator(device='cuda').manual_seed(1)
t = torch.sort(torch.rand(n, device='cuda', generator=g) * 0.04)[0]; t = t - t[0]
if os.environ.get('DIST', 'uniform') == 'edges':      # cfg 5's edge-clustered distribution (bench.py: 24 moving edges)
    k = 24
    e = torch.randint(0, k, (n,), device='cuda', generator=g)
    r = lambda: torch.rand(k, device='cuda', generator=g)
    ex0, ey0, ang, length = r() * Wv, r() * Hv, r() * 3.14159, 60.0 + r() * 200.0
    vx, vy = (r() - 0.5) * 4000.0, (r() - 0.5) * 4000.0
    along = (torch.rand(n, device='cuda', generator=g) - 0.5) * length[e]
    jit = torch.randn(n, device='cuda', generator=g) * 0.7
    x = (ex0[e] + vx[e] * t + along * torch.cos(ang[e]) - jit * torch.sin(ang[e])).clamp_(0, Wv - 1).floor()
    y = (ey0[e] + vy[e] * t + along * torch.sin(ang[e]) + jit * torch.cos(ang[e])).clamp_(0, Hv - 1).floor()
else:
    x = torch.randint(0, Wv, (n,), device='cuda', generator=g).float()
    y = torch.randint(0, Hv, (n,), device='cuda', generator=g).float()
p = torch.randint(0, 2, (n,), device='cuda', generator=g).float() * 2 - 1
grid = torch.empty((bins, Hv, Wv), device='cuda')
st = _lib.stream_ptr()
for i in range(int(os.environ.get('ITERS', '5'))):
    _lib.check(lib.evk_voxelize(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), n, bins, Hv, Wv, _lib.ptr(grid), None, st))
torch.cuda.synchronize()
print('sum', float(grid.sum()), 'abs', float(grid.abs().sum()))
