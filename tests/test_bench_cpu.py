"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys,
and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def _check_reference_line(r, kind):
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'halo_model_spectra_per_sec' and d['unit'] == 'spectra/s'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['steps'] == 1 and d['warmup'] == 0 and d['n_gpus'] == 1 and d['dtype'] == 'f64'
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] == kind and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert d['config'] == json.loads(json.dumps(bench.CONFIG)) and d['config']['workload'].startswith('C3')


def test_reference_arm_prints_one_json_line():
    """With the unmodified reference staged under baseline/_ref (tools/stage_reference.sh) the arm times that
    (kind 'reference'); without it, the oracle port."""
    staged = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'pyhalomodel'))
    _check_reference_line(_run('--impl', 'reference', '--steps', '1', '--warmup', '0'), 'reference' if staged else 'port')


def test_reference_arm_falls_back_to_the_oracle_port():
    env = dict(os.environ, HALOB200_BENCH_PORT='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    _check_reference_line(r, 'port')


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--steps', '1', '--warmup', '0'], cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='only meaningful on a box without a GPU')
def test_product_arm_fails_loudly_without_cuda():
    r = _run('--steps', '1', '--warmup', '0', '--no-cpu', timeout=300)
    assert r.returncode != 0
    assert 'no CPU fallback' in r.stderr
    assert r.stdout.strip() == ''
