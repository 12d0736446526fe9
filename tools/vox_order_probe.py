"""Voxelizer event order probe: grid-stride against chunked CTAs (EVK_VOX_CHUNKED) on the two cfg-5 distributions
(uniform coordinates, events on moving edges), 640x480, CUDA-event timing; checks both orders give the same grid.

Result (profiles/r02c_vox_event_order.jsonl): the order does not matter at 2 M / 4 M events per window (edge-clustered 80.4 vs
80.8 us, uniform 43.2 vs 43.4 us), so the kernel variant was removed again; it is in the history one commit before this note
(`git log -S EVK_VOX_CHUNKED`).  Against the shipped library both rows of a pair time the same grid-stride kernel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evreal_b200 import _lib
lib = _lib.load()
Hv, Wv, bins, n, sets = 480, 640, 5, 4_000_000, 4
g = torch.Generator(device='cuda').manual_seed(1)


def uniform():
    x = torch.randint(0, Wv, (n,), device='cuda', generator=g).float()
    y = torch.randint(0, Hv, (n,), device='cuda', generator=g).float()
    t = torch.sort(torch.rand(n, device='cuda', generator=g) * 0.04)[0]
    p = torch.randint(0, 2, (n,), device='cuda', generator=g).float() * 2 - 1
    return x, y, t - t[0], p


def edges():
    k = 24
    e = torch.randint(0, k, (n,), device='cuda', generator=g)
    t = torch.sort(torch.rand(n, device='cuda', generator=g) * 0.04)[0]
    t = t - t[0]
    r = lambda: torch.rand(k, device='cuda', generator=g)
    ex0, ey0, ang, length = r() * Wv, r() * Hv, r() * 3.14159, 60.0 + r() * 200.0
    vx, vy = (r() - 0.5) * 4000.0, (r() - 0.5) * 4000.0
    along = (torch.rand(n, device='cuda', generator=g) - 0.5) * length[e]
    jit = torch.randn(n, device='cuda', generator=g) * 0.7
    x = (ex0[e] + vx[e] * t + along * torch.cos(ang[e]) - jit * torch.sin(ang[e])).clamp_(0, Wv - 1).floor()
    y = (ey0[e] + vy[e] * t + along * torch.sin(ang[e]) + jit * torch.cos(ang[e])).clamp_(0, Hv - 1).floor()
    p = torch.randint(0, 2, (n,), device='cuda', generator=g).float() * 2 - 1
    return x, y, t, p


grid = torch.empty((bins, Hv, Wv), device='cuda')
st = _lib.stream_ptr()
out = []
for dist, make in (('uniform', uniform), ('edge_clustered', edges)):
    evs = [make() for _ in range(sets)]
    for n_ev in (400_000, 2_000_000, 4_000_000):
        ref = None
        for mode in ('0', '1'):
            os.environ['EVK_VOX_CHUNKED'] = mode
            def run(i):
                x, y, t, p = evs[i % sets]
                _lib.check(lib.evk_voxelize(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), n_ev, bins, Hv, Wv, _lib.ptr(grid), None, st))
            for i in range(4):
                run(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(20):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            run(0)
            got = grid.clone()
            if ref is None:
                ref = got
            diff = float((got - ref).abs().max())
            out.append({'dist': dist, 'events': n_ev, 'chunked': int(mode), 'us': e0.elapsed_time(e1) / 20 * 1e3, 'max_abs_diff_vs_grid_stride': diff})
            print(json.dumps(out[-1]), flush=True)
