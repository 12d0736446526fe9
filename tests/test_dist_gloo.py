"""CPU, world_size 2 over gloo: sequence sharding + the single all-reduce of metric sums reproduce the
single-process dataset means (eval.py:259-266, :367-368)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PER_SEQ = [(10, {'mse': 0.010, 'ssim': 0.80}), (3, {'mse': 0.030, 'ssim': 0.60}), (0, {'mse': -1, 'ssim': -1}),
           (25, {'mse': 0.002, 'ssim': 0.95}), (7, {'mse': 0.050, 'ssim': 0.40})]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from evreal_b200.evaluate import MetricTracker, reduce_metric_sums, shard_sequences
    seqs = [{'name': 'seq%d' % i} for i in range(len(PER_SEQ))]
    mine = shard_sequences(seqs, [float(n + 1) for n, _ in PER_SEQ], world)[rank]
    local = MetricTracker()
    for i in mine:
        n, means = PER_SEQ[i]
        for k, v in means.items():
            local.update(k, v, n)
    red = reduce_metric_sums(local, ['mse', 'ssim'])
    q.put((rank, red.get_average('mse'), red.get_average('ssim'), red.get_count('mse')))
    dist.destroy_process_group()


def test_two_rank_reduction_matches_single_process():
    from evreal_b200.evaluate import MetricTracker
    single = MetricTracker()
    for n, means in PER_SEQ:
        for k, v in means.items():
            single.update(k, v, n)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, mse, ssim, count in got:
        assert count == single.get_count('mse') == 45
        assert abs(mse - single.get_average('mse')) < 1e-15
        assert abs(ssim - single.get_average('ssim')) < 1e-15
