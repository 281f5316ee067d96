"""bench.py contract, CPU tier: the reference arm runs without a GPU and prints ONE JSON line with the agreed keys; the GPU arm
refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run('--impl', 'reference', '--workload', 'small', '--steps', '1', '--warmup', '0')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'Gpixel*slice/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0 and 'workload' in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    r = _run('--impl', 'reference', '--workload', 'small', '--steps', '1', '--warmup', '0', '--gpus', '2', env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_gpu_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('GPU present')
    r = _run('--workload', 'small', '--steps', '1', '--warmup', '1', '--no-cpu')
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
