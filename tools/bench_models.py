"""Throughput of the other BASELINE.json configurations through SequenceBatch (resident events), for profiles/:
cfg 3 FireNet on HQF-shape streams (240x180, MSE+SSIM) and cfg 4 HyperE2VID on MVSEC-shape streams (346x260).
Prints one JSON object per model.  Not the bench contract (bench.py is); same timing rules (CUDA events, warm-up)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--batch', type=int, default=24)
    ap.add_argument('--models', default='firenet,hyper,e2vid')
    args = ap.parse_args()
    import torch
    import evreal_b200 as evk
    from evreal_b200 import synthetic
    from evreal_b200.dataset import MemMapDataset
    from evreal_b200.pipeline import SequenceBatch
    shapes = {'e2vid': (180, 240, 1e6, 24.0), 'firenet': (180, 240, 1e6, 25.0), 'hyper': (260, 346, 5e6, 45.0), 'spade': (180, 240, 1e6, 24.0)}
    for name in args.models.split(','):
        H, W, rate, fps = shapes[name]
        dur = (args.steps + 12) / fps
        dss = [MemMapDataset(synthetic.make_stream(H, W, rate, dur, fps, seed=b), num_bins=5,
                             voxel_method={'method': 'between_frames'}, resident=False) for b in range(args.batch)]
        if name == 'e2vid':
            model = evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, norm_bn=True))
            norm, post = True, 'robust'
        elif name == 'spade':
            model = evk.SpadeE2vid().load_state_dict(synthetic.spade_state_dict(0))
            norm, post = False, 'none'
        elif name == 'hyper':
            model = evk.E2VIDRecurrent(dict(synthetic.HYPER_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, dynamic_decoder=True))
            norm, post = False, 'none'
        else:
            model = evk.FireNet_legacy(dict(synthetic.FIRENET_KWARGS)).load_state_dict(synthetic.firenet_state_dict(0))
            norm, post = True, 'none'
        model.to('cuda')
        batch = SequenceBatch(model, dss, norm, post, resident=True)
        batch.reset()
        batch.wait_uploaded()
        n_items = len(batch)
        idx = 1
        for _ in range(5):
            batch.step(idx, idx % (n_items - 1) + 1, sync=False); idx = idx % (n_items - 1) + 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ev = 0
        for _ in range(args.steps):
            _, _, n = batch.step(idx, idx % (n_items - 1) + 1, sync=False); ev += n; idx = idx % (n_items - 1) + 1
        batch.join()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        batch.finish()
        from evreal_b200.util import normalize_pad
        padded = normalize_pad(batch.voxel, batch.Hp, batch.Wp, norm)
        rows = model.profile_forward(padded)
        rows = model.profile_forward(padded)
        fwd = sum(r[1] for r in rows)
        print(json.dumps({"model": name, "H": H, "W": W, "batch_streams": args.batch, "steps": args.steps,
                          "frames_per_s": args.batch * args.steps / (ms * 1e-3), "events_per_s": ev / (ms * 1e-3),
                          "ms_per_step": ms / args.steps, "forward_ms_eager": fwd, "gflop_per_forward": model.flops_per_forward() / 1e9,
                          "layers": [{"op": r[0], "ms": round(r[1], 4), "tflops": round(r[2] / max(r[1], 1e-9) / 1e9, 1)} for r in rows]}))
        del batch, model
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
