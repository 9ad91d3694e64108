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


def test_cpu_path_under_deadline_runs_the_reference_recipe():
    """bench.py's CPU arm for time-to-converge (the oracle's restatement of the reference's PETSc
    configuration inside the reference's stepper, F evaluated by the reference's own C++-OpenMP
    residual where oracle/_ref is built): on a 16x16 grid it reaches the steady state well inside
    the deadline and reports a speed-up; with a deadline it cannot meet it reports how far it got
    and a lower bound instead."""
    sys.path.insert(0, REPO)
    import bench
    done = bench.cpu_solve_lower_bound(0.05, factor=1.0, min_budget=20.0, max_budget=20.0, nx=16, ny=16)
    assert done['converged'] and done['time_steps_done'] >= 3 and done['newton_iterations_done'] >= 3
    assert done['speedup'] > 0 and 'progress_at_deadline' not in done
    # the same path as the oracle's own driver (which the stepper-trace goldens pin on the reference)
    from oracle import ldc_oracle as orc
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    _, hist = orc.run_simulation(Graph(1.0, 1.0, 16, 16, 1e-2, 1.0, 1.0, bench.RE), None, linear='gmres_ilu0')
    assert len(hist) == done['time_steps_done']
    assert sum(h.iterations for h in hist) == done['newton_iterations_done']
    assert sum(h.function_evaluations for h in hist) == done['residual_evaluations_done']
    cut = bench.cpu_solve_lower_bound(0.05, factor=1.0, min_budget=0.02, max_budget=0.02, nx=48, ny=48)
    assert not cut['converged'] and cut['speedup_greater_than'] > 0
    assert cut['cpu_time_to_converge_s_greater_than'] >= 0.02 and 'time step' in cut['progress_at_deadline']
