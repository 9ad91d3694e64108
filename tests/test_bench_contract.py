"""bench.py contract checks that need no GPU: the reference arm's JSON line (it times the reference's
own C++-OpenMP residual, oracle/_ref, or the oracle port) and the absence of any CPU fallback in the
product arm."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py')] + args, cwd=REPO, env=e,
                          capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("omp", [None, "1"])
def test_reference_arm_prints_the_contract_line(omp):
    """also under OMP_NUM_THREADS=1 (what torchrun exports): the arm must still use every core"""
    r = _run(['--impl', 'reference', '--steps', '2', '--warmup', '1'],
             env={} if omp is None else {'OMP_NUM_THREADS': omp})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'residual_mcells_per_s'
    assert line['unit'] == 'Mcells/s' and line['higher_is_better'] is True and line['dtype'] == 'f64'
    assert line['steps'] == 2 and line['warmup'] == 1 and line['n_gpus'] == 1
    assert line['value'] > 0 and line['ms_per_step'] > 0
    assert abs(line['value'] - 1024 * 1024 / (line['ms_per_step'] * 1e-3) / 1e6) < 1e-6 * line['value']
    cb = line['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['value'] == line['value']
    assert cb['cores'] == (os.cpu_count() or 1)
    assert line['e2e'] == {'value': line['value'], 'unit': 'Mcells/s', 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    assert line['gpu_launches'] == 0 and 'workload' in line['config']


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
             env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    r = _run(['--steps', '1', '--warmup', '1', '--no-solve', '--no-cpu', '--no-tables'])
    assert r.returncode != 0
    assert r.stdout.strip() == ''          # no number without the device
