"""Semantics of the pseudo-time driver (reference: time_stepper.py:9-56), checked
on the CPU with scripted fake solvers -- no device needed."""
import copy

import numpy as np
import pytest

from conftest import make_graph
from lid_driven_cavity_problem_b200.nonlinear_solver.exceptions import SolverDivergedException
from lid_driven_cavity_problem_b200.time_stepper import run_simulation


class ScriptedSolver(object):
    """returns a copy of the graph with U scaled towards a fixed point; diverges on demand"""

    def __init__(self, diverge_when=None, decay=0.5):
        self.calls = []
        self.diverge_when = diverge_when or (lambda graph, n_calls: False)
        self.decay = decay

    def __call__(self, graph):
        self.calls.append(graph.dt)
        if self.diverge_when(graph, len(self.calls)):
            raise SolverDivergedException()
        new = copy.deepcopy(graph)
        new.ns_x_mesh.phi_old = np.array(graph.ns_x_mesh.phi)
        new.ns_x_mesh.phi = self.decay * np.array(graph.ns_x_mesh.phi)
        return new


def _graph(dt=1e-2, bc=10.0):
    return make_graph([1.0, 1.0, 4, 4, dt, 1.0, 1.0, bc])


def test_steady_mode_grows_dt_and_stops_on_the_u_norm():
    solver = ScriptedSolver(decay=0.1)
    out = run_simulation(_graph(), None, solver)
    # |U - U_old|_inf / bc: 0.9/10, 0.09/10, ... < 1e-6 after 6 steps
    assert len(solver.calls) == 6
    assert np.allclose(solver.calls, [1e-2 * 1.2 ** k for k in range(6)])
    assert np.abs(out.ns_x_mesh.phi - out.ns_x_mesh.phi_old).max() / out.bc < 1e-6


def test_divergence_halves_dt_and_retries_the_same_step():
    solver = ScriptedSolver(diverge_when=lambda g, n: g.dt > 3e-3, decay=0.0)
    g = _graph()
    out = run_simulation(g, None, solver, max_steps=1)
    assert solver.calls == [1e-2, 5e-3, 2.5e-3]          # two failures, then success
    assert g.dt == 2.5e-3                                # the caller's graph carries the new dt
    assert np.all(out.ns_x_mesh.phi == 0.0)


def test_gives_up_below_minimum_dt():
    solver = ScriptedSolver(diverge_when=lambda g, n: True)
    with pytest.raises(RuntimeError, match="too low"):
        run_simulation(_graph(dt=4e-6), None, solver, minimum_dt=1e-6)
    assert solver.calls == [4e-6, 2e-6]                  # 1e-6 <= minimum_dt -> give up


def test_fixed_time_mode_clamps_the_last_step_and_doubles_dt_at_the_end():
    solver = ScriptedSolver()
    g = _graph(dt=0.04)
    out = run_simulation(g, 0.1, solver)
    assert np.allclose(solver.calls, [0.04, 0.04, 0.02])  # third step clamped onto final_time
    assert out.dt == pytest.approx(0.04)                  # 0.02 * 2.0 when |final_time - t| <= 1e-5


def test_non_adaptive_divergence_raises_instead_of_looping():
    """the reference retries the same dt forever (time_stepper.py:23-32); this driver raises"""
    solver = ScriptedSolver(diverge_when=lambda g, n: True)
    with pytest.raises(RuntimeError):
        run_simulation(_graph(), None, solver, adaptative_dt=False)
    assert len(solver.calls) == 1


def test_solver_builder_names_on_cpu():
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    assert (solver_builder.CUDA_SOLVER, solver_builder.SCIPY_SOLVER, solver_builder.PETSC_SOLVER) == \
        ('cuda', 'scipy', 'petsc')
    with pytest.raises(AssertionError):
        solver_builder.build('unknown', None)
    with pytest.raises(RuntimeError):
        solver_builder.build('petsc', None)
    # the scipy path needs no device: 8x8 cavity with the oracle residual as residual_f
    from oracle import ldc_oracle as orc
    solve = solver_builder.build('scipy', orc.residual_function)
    g = make_graph([1.0, 1.0, 8, 8, 0.05, 1.0, 1.0, 10.0])
    out = solve(g)
    assert np.array_equal(out.ns_x_mesh.phi_old, g.ns_x_mesh.phi)
    F = orc.residual_function(orc.create_X(out.ns_x_mesh.phi, out.ns_y_mesh.phi, out.pressure_mesh.phi, out), out)
    assert np.abs(F).max() < 1e-6


def _scenarios():
    import importlib.util
    import os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location('generate_golden', os.path.join(GOLDEN, 'generate_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)        # module level only defines helpers; nothing of the reference is touched
    return mod.STEPPER_SCENARIOS


@pytest.mark.parametrize("name", ['steady', 'steady_with_failures', 'steady_fixed_dt', 'fixed_time',
                                  'fixed_time_with_failure', 'fixed_time_exact_multiple', 'give_up'])
def test_replays_the_reference_driver_traces(name, golden):
    """tests/golden/stepper_traces.npz: the REFERENCE's run_simulation (time_stepper.py:9-56) under
    scripted solvers -- every dt a solver call saw, the dt left on the caller's and on the returned
    graph, the final U.  This driver must reproduce them exactly."""
    sc = _scenarios()[name]
    d = golden('stepper_traces.npz')
    calls = []

    def solver(graph):
        calls.append(graph.dt)
        if len(calls) in sc['diverge_on_calls'] or sc['always_diverge']:
            raise SolverDivergedException()
        new = copy.deepcopy(graph)
        new.ns_x_mesh.phi_old = np.array(graph.ns_x_mesh.phi)
        new.ns_x_mesh.phi = sc['decay'] * np.array(graph.ns_x_mesh.phi)
        return new
    g = make_graph([1.0, 1.0, 4, 4, sc['dt'], 1.0, 1.0, 10.0])
    raised = 0
    try:
        out = run_simulation(g, sc['final_time'], solver, minimum_dt=sc['minimum_dt'],
                             adaptative_dt=sc['adaptative_dt'])
    except RuntimeError:
        raised, out = 1, g
    assert raised == int(d[name + '_raised'])
    assert np.array_equal(np.array(calls), d[name + '_calls'])          # exact, not approximate
    assert g.dt == float(d[name + '_caller_dt'])
    assert out.dt == float(d[name + '_out_dt'])
    assert np.array_equal(np.asarray(out.ns_x_mesh.phi, dtype=np.float64), d[name + '_out_U'])
