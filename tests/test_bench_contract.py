"""CPU: the reference arm of bench.py (the one arm that runs without a GPU) prints exactly ONE JSON line with the contract's
keys, and the product arm refuses to run without CUDA instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run('--impl', 'reference', '--steps', '3', '--warmup', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'reconstructed_frames_per_s' and d['unit'] == 'frames/s'
    assert d['higher_is_better'] is True and d['steps'] == 3 and d['value'] > 0 and d['gpu_launches'] == 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and cb['cores'] >= 1 and 'frames' in cb['sample']
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    r = _run('--impl', 'reference', '--steps', '2', '--warmup', '1', env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine WITHOUT a GPU')
def test_product_arm_fails_loudly_without_cuda():
    r = _run('--steps', '1', '--warmup', '1')
    assert r.returncode != 0 and r.stdout.strip() == ''
    assert 'no CUDA device' in r.stderr or 'CUDA' in r.stderr
