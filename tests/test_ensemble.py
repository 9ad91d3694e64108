"""Reynolds ensembles: host-side stepper logic (CPU, incl. a 2-rank gloo run of the sharding
logic) and the batched device solver against one-cavity-at-a-time solves (GPU)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO, make_graph


def test_partition_is_balanced_and_complete():
    from lid_driven_cavity_problem_b200.ensemble import partition
    for n, w in [(512, 8), (512, 1), (10, 4), (3, 8), (0, 2)]:
        parts = [partition(n, w, r) for r in range(w)]
        idx = [i for p in parts for i in range(n)[p]]
        assert idx == list(range(n))
        sizes = [len(range(n)[p]) for p in parts]
        assert max(sizes) - min(sizes) <= 1


def test_stepper_rules_per_cavity():
    """time_stepper.py:9-56 applied independently to every cavity"""
    from lid_driven_cavity_problem_b200.ensemble import EnsembleTimeStepper, RUNNING, STEADY, FAILED
    ts = EnsembleTimeStepper([1e-2, 1e-2, 1e-2, 4e-6], [100.0, 100.0, 1000.0, 10.0])
    dt = ts.begin_step().copy()
    # cavity 0 converges (not steady), 1 diverges, 2 reaches steady state, 3 diverges below minimum_dt
    ts.end_step([3, -3, 4, -5], [1.0, 0.0, 1e-4, 0.0])
    assert np.allclose(ts.t, [1e-2, 0.0, 1e-2, 0.0])
    assert np.allclose(ts.dt, [1.2e-2, 5e-3, 1e-2, 2e-6])
    assert list(ts.status) == [RUNNING, RUNNING, STEADY, RUNNING]
    ts.end_step([3, 3, 0, -3], [1.0, 1.0, 0.0, 0.0])
    assert list(ts.status) == [RUNNING, RUNNING, STEADY, FAILED]     # 1e-6 <= minimum_dt
    assert np.allclose(ts.t[:2], [2.2e-2, 5e-3]) and ts.steps[2] == 1 and ts.retries[1] == 1
    # fixed-time mode clamps the last step
    ts = EnsembleTimeStepper([0.05, 0.08], [10.0, 10.0], final_time=0.1)
    assert np.allclose(ts.begin_step(), [0.05, 0.08])
    ts.end_step([3, 3], [1.0, 1.0])
    assert np.allclose(ts.begin_step(), [0.05, 0.02])
    ts.end_step([3, 3], [1.0, 1.0])
    assert not ts.active().any() and np.allclose(ts.t, 0.1)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["LDC_REPO"])
import numpy as np
import torch.distributed as dist
from lid_driven_cavity_problem_b200 import ensemble
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# fake device solver: cavity k needs k+1 steps to become steady and diverges once at its first step
class FakeSolver(object):
    def __init__(self, graphs, **kw):
        self.graphs, self.batch = graphs, len(graphs)
        self.count = np.zeros(self.batch, int); self.failed_once = np.zeros(self.batch, bool)
        self.newton_iterations = np.zeros(self.batch, int); self.krylov_iterations = np.zeros(self.batch, int)
    def run(self, final_time, minimum_dt, adaptative_dt, max_steps):
        ts = ensemble.EnsembleTimeStepper([g["dt"] for g in self.graphs], [g["bc"] for g in self.graphs])
        while ts.active().any():
            ts.begin_step()
            reasons, du = np.full(self.batch, 3), np.ones(self.batch)
            for b in np.nonzero(ts.active())[0]:
                if not self.failed_once[b]:
                    self.failed_once[b] = True; reasons[b] = -3; continue
                self.count[b] += 1
                if self.count[b] >= self.graphs[b]["need"]: du[b] = 0.0
            ts.end_step(reasons, du)
        return ts
    def fields(self): return None
    def close(self): pass
ensemble.CudaEnsembleSolver = FakeSolver
class G(dict):
    @property
    def bc(self): return self["bc"]
graphs = [G(dt=1e-2, bc=100.0 * (k + 1), need=k + 1) for k in range(7)]
summ, _ = ensemble.run_ensemble(graphs, distributed=True)
assert [s["index"] for s in summ] == list(range(7)), summ
assert [s["steps"] for s in summ] == [k + 1 for k in range(7)], summ
assert all(s["retries"] == 1 and s["status"] == ensemble.STEADY for s in summ)
owners = [s["rank"] for s in summ]
assert owners == [0, 0, 0, 0, 1, 1, 1], owners
# the same with every rank's cavities cut into two concurrent difficulty groups
summ2, _ = ensemble.run_ensemble(graphs, distributed=True, groups=2)
assert [(s["index"], s["steps"], s["retries"], s["rank"]) for s in summ2] == \
       [(s["index"], s["steps"], s["retries"], s["rank"]) for s in summ], summ2
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("rank%dok\n" % rank); sys.stdout.flush()
'''


def test_distributed_sharding_gloo_world2(tmp_path):
    """N>1 host path on CPU: 2 ranks, gloo, batch-sharded ensemble with a fake device solver"""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, LDC_REPO=REPO)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'rank0ok' in r.stdout and 'rank1ok' in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [1, 3])
def test_ensemble_matches_individual_solves(groups):
    """6 cavities, Re 100..3200 (the last one has to halve dt): lock-step batched solve ==
    each cavity solved alone (same kernels, masks must not leak between cavities); also with the
    cavities cut into 3 difficulty groups that run concurrently (own solver, stream and host thread
    each), handed back in the caller's order (the list is deliberately not sorted by Re)"""
    from lid_driven_cavity_problem_b200.ensemble import run_ensemble, STEADY, RUNNING
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    n = 32
    bcs = [1000.0, 100.0, 10000.0, 200.0, 3200.0, 400.0]
    graphs = [make_graph([1.0, 1.0, n, n, 1e-2, 1.0, 1.0, bc]) for bc in bcs]
    summ, (U, V, P) = run_ensemble(graphs, max_steps=6, groups=groups)
    assert [s['bc'] for s in summ] == bcs
    assert len(summ) == 6 and all(s['status'] in (RUNNING, STEADY) for s in summ)
    assert sum(s['retries'] for s in summ) >= 1   # the high-Re cavities have to halve dt
    for b, bc in enumerate(bcs):
        solve = solver_builder.build('cuda', None)
        g = make_graph([1.0, 1.0, n, n, 1e-2, 1.0, 1.0, bc])
        # replay exactly the number of accepted steps the ensemble took for this cavity
        out = run_simulation(g, None, solve, max_steps=summ[b]['steps'])
        assert np.abs(out.ns_x_mesh.phi - U[b]).max() <= 1e-9 * bc, (b, np.abs(out.ns_x_mesh.phi - U[b]).max())
        assert np.abs(out.ns_y_mesh.phi - V[b]).max() <= 1e-9 * bc


@pytest.mark.parametrize("names,final_time", [
    (['steady', 'steady_with_failures', 'give_up'], None),
    (['fixed_time', 'fixed_time_with_failure'], 0.1)])
def test_ensemble_stepper_replays_the_reference_driver_traces(names, final_time, golden):
    """the per-cavity rules of EnsembleTimeStepper against tests/golden/stepper_traces.npz (the
    REFERENCE's run_simulation under scripted solvers): cavities with different scripts advance in
    one batch and each one must see exactly the dt sequence the reference driver produced"""
    import importlib.util
    from conftest import GOLDEN
    from lid_driven_cavity_problem_b200.ensemble import EnsembleTimeStepper, FAILED
    spec = importlib.util.spec_from_file_location('generate_golden', os.path.join(GOLDEN, 'generate_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    scs = [mod.STEPPER_SCENARIOS[n] for n in names]
    d = golden('stepper_traces.npz')
    ts = EnsembleTimeStepper([sc['dt'] for sc in scs], [10.0] * len(scs), final_time=final_time)
    calls = [[] for _ in scs]
    U = np.ones(len(scs))                      # Graph default initial_U = 1.0 (staggered_grid.py:36)
    for _ in range(100):
        act = ts.active()
        if not act.any():
            break
        dt = ts.begin_step()
        reasons, du = np.full(len(scs), 3), np.zeros(len(scs))
        for b in np.nonzero(act)[0]:
            calls[b].append(dt[b])
            if len(calls[b]) in scs[b]['diverge_on_calls'] or scs[b]['always_diverge']:
                reasons[b] = -3
            else:
                du[b] = abs(scs[b]['decay'] * U[b] - U[b])
                U[b] *= scs[b]['decay']
        ts.end_step(reasons, du)
    for b, n in enumerate(names):
        assert np.array_equal(np.array(calls[b]), d[n + '_calls']), n
        assert ts.dt[b] == float(d[n + '_out_dt']), n
        assert (ts.status[b] == FAILED) == bool(d[n + '_raised']), n
