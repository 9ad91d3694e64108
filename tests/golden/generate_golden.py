"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and oracle/_ref):

    python tests/golden/generate_golden.py

It imports the reference's Python package from /root/reference (pure-Python / numpy
residuals, ``_create_X`` / ``_recover_X`` / ``_calculate_jacobian_mask``, the scipy solver
wrapper and the time stepper), the reference's C++ residuals compiled unchanged
(oracle/_ref), and the reference's own golden files, and stores inputs + outputs as small
``.npz`` files.  Nothing here is read at test time except the ``.npz`` outputs;
/root/reference does not exist on the GPU box.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get('LDC_REFERENCE', '/root/reference')


def _import_reference():
    # time_stepper.py:3 and solver_builder import the PETSc wrapper unconditionally;
    # petsc4py is not installable here, so a stub satisfies the import (never called).
    stub = types.ModuleType('petsc4py')
    stub.PETSc = types.SimpleNamespace()
    sys.modules.setdefault('petsc4py', stub)
    # scipy.optimize.minpack (scipy_solver_wrapper.py:4) is gone from recent scipy
    try:
        import scipy.optimize.minpack  # noqa: F401
    except Exception:
        import scipy.optimize as so
        mp = types.ModuleType('scipy.optimize.minpack')
        mp.fsolve = so.fsolve
        sys.modules['scipy.optimize.minpack'] = mp
    sys.path.insert(0, os.path.join(REPO, 'oracle', '_ref'))  # compiled _residual_function
    sys.path.insert(0, REFERENCE)
    import lid_driven_cavity_problem  # noqa: F401
    assert lid_driven_cavity_problem.__file__.startswith(REFERENCE)


# scripted-solver scenarios for the time-stepper traces (also imported by tests/test_time_stepper.py)
STEPPER_SCENARIOS = {
    'steady': dict(dt=1e-2, final_time=None, decay=0.1, diverge_on_calls=(), always_diverge=False,
                   minimum_dt=1e-6, adaptative_dt=True),
    'steady_with_failures': dict(dt=1e-2, final_time=None, decay=0.2, diverge_on_calls=(1, 2, 5, 9),
                                 always_diverge=False, minimum_dt=1e-6, adaptative_dt=True),
    'steady_fixed_dt': dict(dt=1e-2, final_time=None, decay=0.1, diverge_on_calls=(), always_diverge=False,
                            minimum_dt=1e-6, adaptative_dt=False),
    'fixed_time': dict(dt=0.04, final_time=0.1, decay=0.5, diverge_on_calls=(), always_diverge=False,
                       minimum_dt=1e-6, adaptative_dt=True),
    'fixed_time_with_failure': dict(dt=0.03, final_time=0.1, decay=0.5, diverge_on_calls=(2,),
                                    always_diverge=False, minimum_dt=1e-6, adaptative_dt=True),
    'fixed_time_exact_multiple': dict(dt=0.025, final_time=0.1, decay=0.5, diverge_on_calls=(),
                                      always_diverge=False, minimum_dt=1e-6, adaptative_dt=False),
    'give_up': dict(dt=4e-6, final_time=None, decay=0.5, diverge_on_calls=(), always_diverge=True,
                    minimum_dt=1e-6, adaptative_dt=True),
}


def synthetic_state(nx, ny, bc, seed):
    """SURVEY 8(d) synthetic inputs: mixed signs, nothing cancels."""
    rng = np.random.default_rng(seed)
    U = bc * rng.uniform(-1, 1, (nx - 1) * ny)
    V = bc * rng.uniform(-1, 1, nx * (ny - 1))
    P = bc * bc * rng.uniform(-1, 1, nx * ny)
    return U, V, P


def main():
    _import_reference()
    from lid_driven_cavity_problem.staggered_grid import Graph
    from lid_driven_cavity_problem.nonlinear_solver._common import _create_X, _recover_X, \
        _calculate_jacobian_mask
    from lid_driven_cavity_problem.residual_function import pure_python_residual_function, \
        numpy_residual_function, cpp_residual_function, cpp_omp_residual_function

    # ---- 1. the reference's own residual test vector (tests/test_residual_function.py:8-44)
    x = np.array([123.0, 0.002, -0.02] * 9)
    g = Graph(1.0, 1.0, 3, 3, 0.01, 1.0, 1.0, 100.0, initial_P=100.0, initial_U=0.001, initial_V=-0.01)
    F_py = np.array(pure_python_residual_function.residual_function(x, g), dtype=np.float64)
    for f in (numpy_residual_function.residual_function, cpp_residual_function.residual_function,
              cpp_omp_residual_function.residual_function):
        assert np.array_equal(np.asarray(f(x, g)), F_py)
    np.savez(os.path.join(HERE, 'residual_3x3_reference_test.npz'), X=x, F=F_py,
             params=np.array([1.0, 1.0, 3, 3, 0.01, 1.0, 1.0, 100.0, 100.0, 0.001, -0.01]))

    # ---- 2. random states on small ragged grids, all four reference backends bit-equal
    cases = {}
    for k, (nx, ny, bc, dt, sx, sy, rho, mi) in enumerate([
            (3, 2, 10.0, 0.01, 1.0, 1.0, 1.0, 1.0),
            (3, 3, 100.0, 0.01, 1.0, 1.0, 1.0, 1.0),
            (7, 5, 400.0, 0.02, 1.0, 2.0, 1.3, 0.7),
            (16, 9, 1000.0, 0.01, 1.0, 1.0, 1.0, 1.0),
            (33, 20, 1000.0, 0.005, 2.0, 1.0, 0.9, 1.1),
            (64, 64, 3200.0, 0.01, 1.0, 1.0, 1.0, 1.0),
            (100, 37, 1000.0, 0.01, 1.0, 1.0, 1.0, 1.0)]):
        g = Graph(sx, sy, nx, ny, dt, rho, mi, bc)
        U, V, P = synthetic_state(nx, ny, bc, 1000 + nx)
        g.ns_x_mesh.phi_old = 0.9 * U
        g.ns_y_mesh.phi_old = 0.9 * V
        X = _create_X(U, V, P, g)
        # make the dummy slots non-zero too: F must copy them and depend on nothing else
        rng = np.random.default_rng(7 + k)
        Xd = X.copy()
        Xd[1::3][nx - 1::nx] = rng.uniform(-1, 1, ny)
        Xd[2::3][nx * (ny - 1):] = rng.uniform(-1, 1, nx)
        for tag, xx in (('', X), ('_dummy', Xd)):
            F = np.asarray(cpp_residual_function.residual_function(xx, g), dtype=np.float64)
            assert np.array_equal(F, np.asarray(cpp_omp_residual_function.residual_function(xx, g)))
            assert np.array_equal(F, np.asarray(numpy_residual_function.residual_function(xx, g)))
            if nx * ny <= 1000:
                assert np.array_equal(F, np.array(pure_python_residual_function.residual_function(xx, g)))
            cases['X%s_%d' % (tag, k)] = xx
            cases['F%s_%d' % (tag, k)] = F
        cases['U_%d' % k], cases['V_%d' % k], cases['P_%d' % k] = U, V, P
        Ur, Vr, Pr = _recover_X(X, g)
        assert np.array_equal(Ur, U) and np.array_equal(Vr, V) and np.array_equal(Pr, P)
        cases['params_%d' % k] = np.array([sx, sy, nx, ny, dt, rho, mi, bc])
    cases['count'] = np.array(7)
    np.savez_compressed(os.path.join(HERE, 'residual_random_cases.npz'), **cases)

    # ---- 3. Jacobian mask: reference function + the reference's golden files
    masks = {}
    for nx, ny, dof in [(5, 5, 1), (4, 4, 3), (5, 4, 3), (4, 5, 3), (4, 2, 3), (3, 3, 3), (9, 8, 3),
                        (16, 11, 3), (33, 7, 3)]:
        m = _calculate_jacobian_mask(nx, ny, dof)
        assert m.has_sorted_indices and m.indptr.dtype == np.int32 and np.all(m.data == 1.0)
        key = '%d_%d_%d' % (nx, ny, dof)
        masks['indptr_' + key] = m.indptr
        masks['indices_' + key] = m.indices
        gold = os.path.join(REFERENCE, 'tests', 'test_calculate_jacobian_mask', 'J_%s.txt' % key)
        if os.path.exists(gold):
            dense = np.loadtxt(gold)
            assert np.array_equal(dense, m.toarray())
            masks['dense_' + key] = np.packbits(dense.astype(bool), axis=None)
            masks['dense_shape_' + key] = np.array(dense.shape)
    np.savez_compressed(os.path.join(HERE, 'jacobian_masks.npz'), **masks)

    # ---- 4. end-to-end 15x15 fixture (tests/test_run_solver.py:20-52): the shipped U/V files
    #         and a fresh run of the reference scipy path + C++ residual
    from lid_driven_cavity_problem.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem.time_stepper import run_simulation
    exp_U = np.loadtxt(os.path.join(REFERENCE, 'tests', 'test_run_solver', 'U.txt'))
    exp_V = np.loadtxt(os.path.join(REFERENCE, 'tests', 'test_run_solver', 'V.txt'))
    solver = solver_builder.build('scipy', cpp_residual_function.residual_function)
    g = Graph(1.0, 1.0, 15, 15, 0.05, 1.0, 1.0, 10.0)
    res = run_simulation(g, 0.1, solver)
    U = np.array(res.ns_x_mesh.phi)
    V = np.array(res.ns_y_mesh.phi)
    print('15x15 scipy rerun vs shipped fixture: max|dU|=%.3e max|dV|=%.3e'
          % (np.abs(U - exp_U).max(), np.abs(V - exp_V).max()))
    assert np.allclose(U, exp_U, rtol=1e-4, atol=1e-1) and np.allclose(V, exp_V, rtol=1e-4, atol=1e-1)
    np.savez(os.path.join(HERE, 'small_case_15x15.npz'), U_shipped=exp_U, V_shipped=exp_V,
             U_scipy_rerun=U, V_scipy_rerun=V, P_scipy_rerun=np.array(res.pressure_mesh.phi),
             params=np.array([1.0, 1.0, 15, 15, 0.05, 1.0, 1.0, 10.0, 0.1]))

    # ---- 5. Ghia-Ghia-Shin centre-line tables shipped with the reference (Re=100, 400)
    gU = np.loadtxt(os.path.join(REFERENCE, 'experiments', 'ghia_ghia_shin_results', 'U.txt'))
    gV = np.loadtxt(os.path.join(REFERENCE, 'experiments', 'ghia_ghia_shin_results', 'V.txt'))
    np.savez(os.path.join(HERE, 'ghia_reference_tables.npz'), U=gU, V=gV)

    # ---- 6. central differencing: only the reference's numba backend honours graph.use_cds
    #         (numba_residual_function.py:128-137,202-211,247); with use_cds=False it must equal
    #         the C++ residual, with use_cds=True it is the golden vector of the CDS switch
    from lid_driven_cavity_problem.residual_function import numba_residual_function
    cds = {}
    cds_cases = [(3, 2, 10.0, 0.01, 1.0, 1.0, 1.0, 1.0), (7, 5, 400.0, 0.02, 1.0, 2.0, 1.3, 0.7),
                 (33, 20, 1000.0, 0.005, 2.0, 1.0, 0.9, 1.1), (64, 48, 3200.0, 0.01, 1.0, 1.0, 1.0, 1.0)]
    for k, (nx, ny, bc, dt, sx, sy, rho, mi) in enumerate(cds_cases):
        g = Graph(sx, sy, nx, ny, dt, rho, mi, bc)
        U, V, P = synthetic_state(nx, ny, bc, 2000 + nx)
        g.ns_x_mesh.phi_old = 0.9 * U
        g.ns_y_mesh.phi_old = 0.9 * V
        X = _create_X(U, V, P, g)
        g.use_cds = False
        F_uds = np.asarray(numba_residual_function.residual_function(X, g), dtype=np.float64)
        assert np.array_equal(F_uds, np.asarray(cpp_residual_function.residual_function(X, g)))
        g.use_cds = True
        F_cds = np.asarray(numba_residual_function.residual_function(X, g), dtype=np.float64)
        assert not np.array_equal(F_cds, F_uds)
        cds['X_%d' % k], cds['F_cds_%d' % k], cds['F_uds_%d' % k] = X, F_cds, F_uds
        cds['U_old_%d' % k], cds['V_old_%d' % k] = 0.9 * U, 0.9 * V
        cds['params_%d' % k] = np.array([sx, sy, nx, ny, dt, rho, mi, bc])
    cds['count'] = np.array(len(cds_cases))
    np.savez_compressed(os.path.join(HERE, 'residual_cds_cases.npz'), **cds)

    # ---- 7. traces of the reference pseudo-time driver (time_stepper.py:9-56) under scripted
    #         solvers: which dt every solver call saw, the dt left on the caller's graph and on
    #         the returned graph, the final U.  tests/test_time_stepper.py replays them.
    import copy
    from lid_driven_cavity_problem.nonlinear_solver.exceptions import SolverDivergedException
    traces = {}
    for name, sc in STEPPER_SCENARIOS.items():
        calls = []

        def solver(graph, sc=sc, calls=calls):
            calls.append(graph.dt)
            if len(calls) in sc['diverge_on_calls'] or sc['always_diverge']:
                raise SolverDivergedException()
            new = copy.deepcopy(graph)
            new.ns_x_mesh.phi_old = np.array(graph.ns_x_mesh.phi)
            new.ns_x_mesh.phi = sc['decay'] * np.array(graph.ns_x_mesh.phi)
            return new
        g = Graph(1.0, 1.0, 4, 4, sc['dt'], 1.0, 1.0, 10.0)
        raised = 0
        try:
            out = run_simulation(g, sc['final_time'], solver, minimum_dt=sc['minimum_dt'],
                                 adaptative_dt=sc['adaptative_dt'])
        except RuntimeError:
            raised, out = 1, g
        traces[name + '_calls'] = np.array(calls)
        traces[name + '_caller_dt'] = np.array(g.dt)
        traces[name + '_out_dt'] = np.array(out.dt)
        traces[name + '_out_U'] = np.array(out.ns_x_mesh.phi, dtype=np.float64)
        traces[name + '_raised'] = np.array(raised)
    np.savez(os.path.join(HERE, 'stepper_traces.npz'), **traces)
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
