"""Minimal driver for ncu: a few lock-step frames of bench.py's workload (E2VID, ECD-shape streams) and nothing else.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --steps 3
    ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 26 -c 3 -o gpurun_out/prof_conv \
        python tools/profile_step.py --steps 3
Numbers printed under a profiler are never bench values.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--batch', type=int, default=24)
    ap.add_argument('--model', default='e2vid', choices=['e2vid', 'firenet', 'hyper'])
    ap.add_argument('--voxel-only', action='store_true', help='cfg 5 voxelizer launches only (640x480, 4M events)')
    ap.add_argument('--voxel-events', type=int, default=4_000_000)
    ap.add_argument('--lpips', default='', help="'lpips' or 'lpips-vgg': batched LPIPS calls only (36 pairs at 240x180)")
    args = ap.parse_args()
    import torch
    import evreal_b200 as evk
    from evreal_b200 import _lib, synthetic
    from evreal_b200.dataset import MemMapDataset
    from evreal_b200.pipeline import SequenceBatch
    if args.voxel_only:
        lib = _lib.load()
        n, Hv, Wv = args.voxel_events, 480, 640
        g = torch.Generator(device='cuda').manual_seed(1)
        x = torch.randint(0, Wv, (n,), device='cuda', generator=g).float()
        y = torch.randint(0, Hv, (n,), device='cuda', generator=g).float()
        t = torch.sort(torch.rand(n, device='cuda', generator=g) * 0.04)[0]
        p = torch.randint(0, 2, (n,), device='cuda', generator=g).float() * 2 - 1
        grid = torch.empty((5, Hv, Wv), dtype=torch.float32, device='cuda')
        for _ in range(args.steps):
            _lib.check(lib.evk_voxelize(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), n, 5, Hv, Wv, _lib.ptr(grid), None,
                                        _lib.stream_ptr()))
        torch.cuda.synchronize()
        return
    if args.lpips:
        sys.path.insert(0, ROOT)
        import bench
        from evreal_b200.lpips import LpipsNet
        ln = LpipsNet(args.lpips, bench.seeded_lpips_weights('alex' if args.lpips == 'lpips' else 'vgg'), 180, 240, batch=36)
        a, b = torch.rand((36, 180, 240), device='cuda'), torch.rand((36, 180, 240), device='cuda')
        for _ in range(args.steps):
            ln(a, b)
        torch.cuda.synchronize()
        return
    shapes = {'e2vid': (180, 240, 1e6, 24.0), 'firenet': (180, 240, 1e6, 25.0), 'hyper': (260, 346, 5e6, 45.0)}
    H, W, rate, fps = shapes[args.model]
    dur = (args.steps + 4) / fps
    dss = []
    for b in range(args.batch):
        a = synthetic.make_stream(H, W, rate, dur, fps, seed=b)
        dss.append(MemMapDataset(a, num_bins=5, voxel_method={'method': 'between_frames'}, resident=False))
    if args.model == 'e2vid':
        model = evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, norm_bn=True))
        norm, post = True, 'robust'
    elif args.model == 'hyper':
        model = evk.E2VIDRecurrent(dict(synthetic.HYPER_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, dynamic_decoder=True))
        norm, post = False, 'none'
    else:
        model = evk.FireNet_legacy(dict(synthetic.FIRENET_KWARGS)).load_state_dict(synthetic.firenet_state_dict(0))
        norm, post = True, 'none'
    model.to('cuda')
    batch = SequenceBatch(model, dss, norm, post, resident=True)
    batch.reset()
    for i in range(args.steps):
        batch.step(1 + i)
    torch.cuda.synchronize()
    print('profiled', args.steps, 'steps; launches per step', batch.launches)


if __name__ == '__main__':
    main()
