"""CPU: host-side runtime pieces that need no GPU -- the asynchronous writer and the catch-and-continue /
rank-safe reduce structure of evaluate() (eval.py:344-375; ADVICE round 1: a failing rank must not leave the others
waiting in the all-reduce)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_async_writer_keeps_order_and_surfaces_errors(tmp_path):
    from evreal_b200.eval_utils import AsyncWriter, append_result, append_timestamp, set_writer, truncate_file
    w = AsyncWriter()
    old = set_writer(w)
    try:
        p = str(tmp_path / 'mse.txt')
        truncate_file(p)
        for i in range(200):
            append_result(p, i, i / 7.0)
        append_result(p, [200, 201], [1.0, 2.0])
        append_timestamp(str(tmp_path / 'timestamps.txt'), 3, 0.125)
        w.flush()
        rows = open(p).read().splitlines()
        assert rows[:3] == ['0 0.00000', '1 0.14286', '2 0.28571'] and rows[-2:] == ['200 1.00000', '201 2.00000'] and len(rows) == 202
        assert open(tmp_path / 'timestamps.txt').read() == '3 0.125000000000000\n'
        append_result(str(tmp_path / 'no_such_dir' / 'x.txt'), 0, 1.0)
        with pytest.raises(OSError):
            w.flush()
        append_result(p, 202, 3.0)                 # the writer keeps working after a reported error
        w.flush()
        assert open(p).read().splitlines()[-1] == '202 3.00000'
    finally:
        set_writer(old)
        w.close()


def test_synchronous_writers_are_unchanged_without_a_writer(tmp_path):
    from evreal_b200.eval_utils import append_result
    p = str(tmp_path / 'a.txt')
    append_result(p, 5, 0.123456)
    append_result(p, 7, 3, is_int=True)
    assert open(p).read() == '5 0.12346\n7 3\n'


def _worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from evreal_b200 import evaluate as ev

    class FakeModel:
        num_encoders = 0

    def fake_model(model_name, path, device=None):
        if model_name == 'BROKEN':
            raise RuntimeError("checkpoint unreadable")
        return FakeModel()

    def fake_dataset(eval_config, method_name, model, method_config, dataset_config, metrics, output_root, write_files, rank_, world_,
                     lockstep, lpips_weights, local, writer):
        # rank 1 finishes one sequence of dataset B and then fails; rank 0 finishes two of every dataset
        if dataset_config['name'] == 'B' and rank_ == 1:
            local.update('mse', 0.5, 10)
            raise IndexError("3 events are out of bounds")
        local.update('mse', 0.1 * (rank_ + 1), 20)
        local.update('mse', 0.3, 5)

    ev.get_model_from_checkpoint_path = fake_model
    ev._eval_dataset = fake_dataset
    ev.get_method_config = lambda name, root: {'model_name': name, 'model_path': 'x'}
    ev.get_eval_configs = lambda names, root: [{'name': n, 'histeq': 'none'} for n in names]
    ev.get_dataset_configs = lambda names, root: [{'name': n} for n in names]
    res = ev.evaluate(['BROKEN', 'OK'], ['std'], ['A', 'B'], ['mse'], write_files=False, rank=rank, world_size=world)
    out = {m: {d: (tr.get_count('mse'), tr.get_average('mse')) for d, tr in per.items()} for m, per in res['std'].items()}
    q.put((rank, out, ev.last_timings['failures']))
    dist.barrier()
    dist.destroy_process_group()


def test_failing_rank_still_joins_every_reduce():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, None, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, out, failures = q.get(timeout=120)
        got[rank] = (out, failures)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] == got[1][0]                          # identical results on every rank
    out = got[0][0]
    assert out['BROKEN'] == {'A': (0, 0.0), 'B': (0, 0.0)}  # method skipped everywhere, collectives still matched
    n, mean = out['OK']['A']
    assert n == 50 and abs(mean - (0.1 * 20 + 0.3 * 5 + 0.2 * 20 + 0.3 * 5) / 50) < 1e-12
    n, mean = out['OK']['B']                               # rank 1 contributed the sequence it finished before failing
    assert n == 35 and abs(mean - (0.1 * 20 + 0.3 * 5 + 0.5 * 10) / 35) < 1e-12
    assert got[0][1] == 1 and got[1][1] == 2               # rank 0: model failure; rank 1: model + dataset failure


def test_command_line_flags_reach_evaluate(monkeypatch, capsys):
    """python -m evreal_b200.evaluate keeps eval.py's flags (-c / -m / -d / -qm, eval.py:447-455) and hands them to
    evaluate() unchanged; lock-step runs do not write per-frame files (the lock-step form has no per-sequence tracker)."""
    from evreal_b200 import evaluate as ev
    seen = {}

    def fake(methods, configs, datasets, metrics, **kw):
        seen.update(methods=methods, configs=configs, datasets=datasets, metrics=metrics, **kw)
        tr = ev.MetricTracker()
        tr.update('mse', 0.25, 4)
        return {'std': {'E2VID': {'ECD': tr}}}

    monkeypatch.setattr(ev, 'evaluate', fake)
    monkeypatch.delenv('RANK', raising=False)
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    ev.main(['-m', 'E2VID', 'FireNet', '-c', 'std', '-d', 'ECD', 'HQF', '-qm', 'mse', 'ssim', 'lpips', '--lockstep', '16'])
    assert seen['methods'] == ['E2VID', 'FireNet'] and seen['configs'] == ['std'] and seen['datasets'] == ['ECD', 'HQF']
    assert seen['metrics'] == ['mse', 'ssim', 'lpips'] and seen['lockstep'] == 16 and seen['write_files'] is False
    assert seen['rank'] == 0 and seen['world_size'] == 1 and seen['config_root'] == 'config'
    assert "std / E2VID / ECD: {'mse': 0.25}" in capsys.readouterr().out
    ev.main(['-m', 'E2VID', '-d', 'ECD'])
    assert seen['configs'] is None and seen['metrics'] is None and seen['write_files'] is True and seen['lockstep'] == 0
    with pytest.raises(SystemExit):
        ev.main(['-c', 'std'])                     # -m is required
