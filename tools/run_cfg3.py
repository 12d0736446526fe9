"""BASELINE cfg 3 through the reference-facing plugin surface: FireNet on 16 synthetic HQF-shape sequences (240x180, 5 s,
rate U(0.5, 2) Mev/s per sequence, 25 Hz frames), driven by config/{method,eval,dataset}/*.json and a checkpoint
file exactly like EVREAL's eval.py, sequences sharded across the ranks of one torchrun job (evaluate.shard_sequences) and
the dataset means aggregated by ONE all-reduce (evaluate.reduce_metric_sums, NCCL).

    python tools/run_cfg3.py --root /tmp/cfg3                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tools/run_cfg3.py --root /tmp/cfg3                         # 2 GPUs: same means, sequences split 8 / 8
Prints one JSON line on rank 0.  Not the bench contract (bench.py is); sequences run one at a time with batch 1, like the
reference.  Weights are seeded random (the real checkpoint does not exist on the GPU box); the shapes are FireNet's."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_tree(root, n_seq, duration):
    import torch
    from evreal_b200 import synthetic
    H, W, _, _, fps = synthetic.SHAPES['hqf']
    for d in ('config/method', 'config/eval', 'config/dataset', 'pretrained/FireNet', 'data/HQF16'):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    ck = os.path.join(root, 'pretrained/FireNet/model.pth')
    torch.save({'config': {'model': dict(synthetic.FIRENET_KWARGS)}, 'state_dict': synthetic.firenet_state_dict(0)}, ck)
    json.dump({'model_name': 'FireNet', 'model_path': ck, 'event_tensor_normalization': True, 'post_process_norm': 'none'},
              open(os.path.join(root, 'config/method/FireNet.json'), 'w'))
    json.dump({'save_images': False, 'histeq': 'none', 'eval_infer_all': False, 'ts_tol_ms': 1.0, 'create_video': False,
               'dataset_kwargs': {'num_bins': 5, 'voxel_method': {'method': 'between_frames'}}},
              open(os.path.join(root, 'config/eval/std.json'), 'w'))
    seqs = {}
    for i in range(n_seq):
        synthetic.write_sequence(os.path.join(root, 'data/HQF16', 'seq%02d' % i), H, W, synthetic.hqf_rate(i), duration, fps, seed=i)
        seqs['seq%02d' % i] = {'start_time_s': 0.0, 'end_time_s': duration}
    json.dump({'root_path': os.path.join(root, 'data/HQF16'), 'sequences': seqs},
              open(os.path.join(root, 'config/dataset/HQF16.json'), 'w'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--root', default='/tmp/evk_cfg3')
    ap.add_argument('--sequences', type=int, default=16)
    ap.add_argument('--duration', type=float, default=5.0)
    ap.add_argument('--lockstep', type=int, default=0, help='sequences per lock-step batch on every rank (0: one at a time, like the reference)')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    torch.cuda.set_device(local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                                   # NCCL's banner goes to fd 1
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    if rank == 0 and not os.path.exists(os.path.join(args.root, 'config/dataset/HQF16.json')):
        write_tree(args.root, args.sequences, args.duration)
    if world > 1:
        dist.barrier()
    from evreal_b200 import evaluate as ev
    cfg_root = os.path.join(args.root, 'config')
    ev.evaluate(['FireNet'], ['std'], ['HQF16'], ['mse', 'ssim'], config_root=cfg_root, write_files=False, rank=rank, world_size=world, lockstep=args.lockstep)  # warm-up (page cache, plans)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = ev.evaluate(['FireNet'], ['std'], ['HQF16'], ['mse', 'ssim'], config_root=cfg_root, write_files=False, rank=rank, world_size=world, lockstep=args.lockstep)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    tr = res['std']['FireNet']['HQF16']
    phases = dict(ev.last_timings)
    if world > 1:                                   # slowest rank per phase
        keys = sorted(phases)
        v = torch.tensor([phases[k] for k in keys], dtype=torch.float64, device='cuda')
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        phases = {k: float(x) for k, x in zip(keys, v.cpu())}
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        os.dup2(saved, 1)
        n = tr.get_count('mse')
        print(json.dumps({'config': 'cfg3: FireNet, %d HQF-shape sequences (240x180, %.0f s, U(0.5,2) Mev/s, 25 Hz), between_frames, MSE+SSIM' % (args.sequences, args.duration),
                          'n_gpus': world, 'lockstep': args.lockstep, 'frames_evaluated': n, 'seconds': dt, 'frames_per_s': n / dt,
                          'phase_seconds_max_over_ranks': phases, 'loop_frames_per_s': n / max(phases['loop_s'], 1e-9),
                          'mse_mean': repr(tr.get_average('mse')), 'ssim_mean': repr(tr.get_average('ssim')),
                          'api': 'evreal_b200.evaluate.evaluate (config/*.json + checkpoint, batch 1 per sequence, one all-reduce of [sum(score*n), sum(n)])'}))


if __name__ == '__main__':
    main()
