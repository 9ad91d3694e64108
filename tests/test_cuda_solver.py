"""GPU tests of the device-resident Newton-Krylov solver and the time stepper, against the CPU
oracle (independent sparse-direct Newton) and the reference's own end-to-end fixture."""
import ctypes
from copy import deepcopy

import numpy as np
import pytest

from conftest import make_graph

pytestmark = pytest.mark.gpu

TIGHT = dict(snes_rtol=1e-13, snes_atol=1e-9, snes_stol=1e-15, ksp_rtol=1e-8, snes_max_it=30, gmres_refine=1)


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _developed_graph(nx, ny, bc, steps=3, dt=1e-2):
    """a non-trivial smooth state: a few oracle pseudo-time steps from the reference's start"""
    from oracle import ldc_oracle as orc
    g = make_graph([1.0, 1.0, nx, ny, dt, 1.0, 1.0, bc])
    g, _ = orc.run_simulation(g, None, linear='direct', max_steps=steps)
    return g


def test_linear_solve_vs_direct():
    """FGMRES + multigrid on the device FD Jacobian == sparse direct solve of the oracle's"""
    import torch
    from oracle import ldc_oracle as orc
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.lib()
    for nx, ny, pc in [(32, 32, _native.PC_MG_VANKA), (48, 40, _native.PC_MG_VANKA), (15, 15, _native.PC_VANKA)]:
        g = _developed_graph(nx, ny, 100.0)
        X = orc.create_X(g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi, g)
        g.ns_x_mesh.phi_old = g.ns_x_mesh.phi.copy()
        g.ns_y_mesh.phi_old = g.ns_y_mesh.phi.copy()
        g.dt = 0.05
        F = orc.residual_function(X, g)
        mask = orc.jacobian_mask_arrays(nx, ny, 3)
        J = orc.fd_jacobian(X, g, F0=F, mask=mask)
        y_ref = orc._direct_solve(mask, J, F)
        opt = _native.LdcSolverOptions()
        lib.ldc_solver_default_options(ctypes.byref(opt))
        opt.ksp_rtol, opt.pc_type, opt.ksp_max_it = 1e-8, pc, 3000
        grid = _native.grid_from_graph(g)
        h = ctypes.c_void_p()
        _native.check(lib.ldc_solver_create(ctypes.byref(grid), ctypes.byref(opt), ctypes.byref(h)))
        try:
            st = _native.current_stream()
            Xd, Fd = torch.from_numpy(X).cuda(), torch.from_numpy(F).cuda()
            yd = torch.empty_like(Xd)
            _native.check(lib.ldc_solver_set_state(h, _native.ptr(Xd), st))
            its, reason, rel = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
            _native.check(lib.ldc_solver_linear_solve(h, _native.ptr(Fd), _native.ptr(yd), ctypes.byref(its),
                                                      ctypes.byref(reason), ctypes.byref(rel), st))
            assert reason.value in (2, 3), (nx, ny, reason.value, its.value, rel.value)
            y = yd.cpu().numpy()
            # true residual of the device solution on the oracle's matrix
            r = F - orc.spmv(mask, J, y)
            assert np.linalg.norm(r) <= 2e-8 * np.linalg.norm(F)
            # solutions agree up to the pressure constant (the Jacobian's null space)
            dP = y[0::3] - y_ref[0::3]
            assert _rel(y[1::3], y_ref[1::3]) < 1e-5 and _rel(y[2::3], y_ref[2::3]) < 1e-5
            assert (dP - dP.mean()).std() < 1e-5 * np.abs(y_ref[0::3] - y_ref[0::3].mean()).max()
            if pc == _native.PC_MG_VANKA:
                assert lib.ldc_solver_num_levels(h) >= 3
                assert its.value < 80, its.value
        finally:
            lib.ldc_solver_destroy(h)


@pytest.mark.parametrize("nx,ny,bc", [(32, 32, 100.0), (64, 64, 1000.0), (40, 24, 400.0), (15, 15, 10.0),
                                      (33, 20, 100.0), (100, 38, 400.0), (3, 2, 10.0)])
def test_one_step_converged_state_parity(nx, ny, bc):
    """north_star: converged U/V/P agree to 1e-8 relative (P up to its constant)."""
    from oracle import ldc_oracle as orc
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    g0 = _developed_graph(nx, ny, bc, steps=2)
    g0.dt = 0.02
    ref, st = orc.solve_step(deepcopy(g0), linear='direct', rtol=1e-13, atol=1e-9, stol=1e-15, max_it=30)
    solve = solver_builder.build('cuda', None, **TIGHT)
    U_in = np.array(g0.ns_x_mesh.phi)
    out = solve(g0)
    assert np.array_equal(g0.ns_x_mesh.phi, U_in)                     # input not modified
    assert np.array_equal(out.ns_x_mesh.phi_old, g0.ns_x_mesh.phi)    # phi_old <- phi
    assert np.array_equal(out.pressure_mesh.phi_old, g0.pressure_mesh.phi)
    assert _rel(out.ns_x_mesh.phi, ref.ns_x_mesh.phi) < 1e-8
    assert _rel(out.ns_y_mesh.phi, ref.ns_y_mesh.phi) < 1e-8
    P, Pr = out.pressure_mesh.phi, ref.pressure_mesh.phi
    assert np.abs((P - P.mean()) - (Pr - Pr.mean())).max() < 1e-8 * np.abs(Pr - Pr.mean()).max()
    assert abs(out.steady_state_norm() -
               np.abs(out.ns_x_mesh.phi - out.ns_x_mesh.phi_old).max() / bc) < 1e-12


@pytest.mark.parametrize("solver_name", ["cuda", "scipy"])
def test_small_case_reference_fixture(solver_name, golden):
    """The reference's end-to-end regression (tests/test_run_solver.py:20-52): 15x15, Re=10,
    dt=0.05, final_time=0.1, compared with its shipped U/V at its own tolerance."""
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    d = golden('small_case_15x15.npz')
    solver = solver_builder.build(solver_name, cuda_residual_function.residual_function)
    graph = Graph(1.0, 1.0, 15, 15, 0.05, 1.0, 1.0, (1.0 * 10) / (1.0 * 1.0))
    result = run_simulation(graph, 0.1, solver)
    U, V = np.array(result.ns_x_mesh.phi), np.array(result.ns_y_mesh.phi)
    assert np.allclose(U, d['U_shipped'], rtol=1e-4, atol=1e-1)
    assert np.allclose(V, d['V_shipped'], rtol=1e-4, atol=1e-1)
    if solver_name == 'scipy':   # same algorithm as the fixture's generator: tight agreement
        assert np.abs(U - d['U_shipped']).max() < 1e-8 and np.abs(V - d['V_shipped']).max() < 1e-8


def test_small_case_tight():
    """same case, tight tolerances: the cuda solve lands on the fixture to ~1e-8"""
    import numpy as np
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    d = np.load(__import__('os').path.join(__import__('conftest').GOLDEN, 'small_case_15x15.npz'))
    solver = solver_builder.build('cuda', None, **TIGHT)
    result = run_simulation(Graph(1.0, 1.0, 15, 15, 0.05, 1.0, 1.0, 10.0), 0.1, solver)
    assert np.abs(result.ns_x_mesh.phi - d['U_shipped']).max() < 1e-7
    assert np.abs(result.ns_y_mesh.phi - d['V_shipped']).max() < 1e-7


def test_steady_state_matches_oracle_and_ghia():
    """steady 32x32 Re=100 with the reference recipe (P=U=V=1 start, dt=1e-2, x1.2 growth)"""
    from oracle import ldc_oracle as orc
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    n, bc = 32, 100.0
    ref, hist = orc.run_simulation(make_graph([1.0, 1.0, n, n, 1e-2, 1.0, 1.0, bc]), None, linear='direct')
    steps = []
    out = run_simulation(make_graph([1.0, 1.0, n, n, 1e-2, 1.0, 1.0, bc]), None,
                         solver_builder.build('cuda', None), on_step=lambda g, t: steps.append(t))
    assert abs(len(steps) - len(hist)) <= 2
    # both stop by the same loose steady test (1e-6 * bc), so they agree to about that level
    assert np.abs(out.ns_x_mesh.phi - ref.ns_x_mesh.phi).max() < 5e-4 * bc
    assert np.abs(out.ns_y_mesh.phi - ref.ns_y_mesh.phi).max() < 5e-4 * bc


def test_divergence_halves_dt_and_recovers():
    """64x64 Re=3200 diverges at dt=1e-2 (so does the reference path); the stepper halves dt"""
    from lid_driven_cavity_problem_b200 import options
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem_b200.nonlinear_solver.exceptions import SolverDivergedException
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    g = make_graph([1.0, 1.0, 64, 64, 1e-2, 1.0, 1.0, 3200.0])
    solve = solver_builder.build('cuda', None)
    with pytest.raises(SolverDivergedException):
        solve(g)
    assert g.dt == 1e-2
    out = run_simulation(g, None, solve, max_steps=1)
    assert g.dt < 1e-2 and np.isfinite(out.ns_x_mesh.phi).all()
    assert out.newton_report.reason > 0
    options.IGNORE_DIVERGED = True
    try:
        res = solve(make_graph([1.0, 1.0, 64, 64, 1e-2, 1.0, 1.0, 3200.0]))
        assert res.newton_report.reason <= 0
    finally:
        options.IGNORE_DIVERGED = False
    with pytest.raises(RuntimeError):
        run_simulation(make_graph([1.0, 1.0, 64, 64, 1e-2, 1.0, 1.0, 3200.0]), None, solve,
                       adaptative_dt=False)


def test_solver_builder_contract():
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    with pytest.raises(AssertionError):
        solver_builder.build('nope', None)
    with pytest.raises(TypeError):
        solver_builder.build('cuda', None, no_such_option=1)(make_graph([1.0, 1.0, 8, 8, 1e-2, 1.0, 1.0, 10.0]))


def test_converged_state_parity_baseline_grid_256():
    """north_star tolerance (converged U/V/P to 1e-8 relative) on a BASELINE.json grid: 256x256
    Re=1000, one pseudo-time step with tight tolerances from a developed state.  A full
    independent Newton needs ~5 sparse LU factorisations of 196 608 unknowns (~30 s each), so
    the check is a posteriori: ONE direct Newton correction of the oracle (its FD Jacobian, its
    residual, sparse LU) at the device solution.  Newton converges quadratically, so that
    correction is the distance to the oracle's converged state."""
    from oracle import ldc_oracle as orc
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    n, bc = 256, 1000.0
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    # a developed state: two accepted steps of the reference recipe (dt is halved where Newton diverges)
    g = run_simulation(make_graph([1.0, 1.0, n, n, 1e-2, 1.0, 1.0, bc]), None, solver_builder.build('cuda', None),
                       max_steps=2)
    dt = g.dt / 2.0
    g0 = make_graph([1.0, 1.0, n, n, dt, 1.0, 1.0, bc], None, g.ns_x_mesh.phi, g.ns_y_mesh.phi,
                    g.pressure_mesh.phi)
    out = solver_builder.build('cuda', None, **TIGHT)(g0)
    assert out.newton_report.reason > 0
    # the oracle's view of that step: phi_old = the entry state, X = the device solution
    chk = make_graph([1.0, 1.0, n, n, dt, 1.0, 1.0, bc], None, out.ns_x_mesh.phi, out.ns_y_mesh.phi,
                     out.pressure_mesh.phi)
    chk.ns_x_mesh.phi_old = np.array(g0.ns_x_mesh.phi)
    chk.ns_y_mesh.phi_old = np.array(g0.ns_y_mesh.phi)
    X = orc.create_X(chk.ns_x_mesh.phi, chk.ns_y_mesh.phi, chk.pressure_mesh.phi, chk)
    F = orc.residual_function(X, chk)
    mask = orc.jacobian_mask_arrays(n, n, 3)
    J = orc.fd_jacobian(X, chk, F0=F, mask=mask)
    delta = orc._direct_solve(mask, J, F)    # X_oracle = X - delta (+ quadratically small terms)
    U, V, P = out.ns_x_mesh.phi, out.ns_y_mesh.phi, out.pressure_mesh.phi
    dP = delta[0::3] - delta[0::3].mean()
    assert np.abs(delta[1::3]).max() < 1e-8 * np.abs(U).max()
    assert np.abs(delta[2::3]).max() < 1e-8 * np.abs(V).max()
    assert np.abs(dP).max() < 1e-8 * np.abs(P - P.mean()).max()


def test_ignore_diverged_returns_the_diverged_iterate():
    """options.IGNORE_DIVERGED: the reference hands back whatever SNES ended with
    (petsc_solver_wrapper.py:90-103).  The returned fields must therefore differ from the entry
    state -- a rolled-back state would make run_simulation see max|U-U_old| = 0 and report
    'steady' after a failed step."""
    from lid_driven_cavity_problem_b200 import options
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    solve = solver_builder.build('cuda', None)
    g = make_graph([1.0, 1.0, 64, 64, 1e-2, 1.0, 1.0, 3200.0])
    options.IGNORE_DIVERGED = True
    try:
        res = solve(g)
    finally:
        options.IGNORE_DIVERGED = False
    assert res.newton_report.reason <= 0
    assert res.steady_state_norm() > 1e-6
    assert not np.array_equal(res.ns_x_mesh.phi, res.ns_x_mesh.phi_old)
    assert np.array_equal(res.ns_x_mesh.phi_old, g.ns_x_mesh.phi)
    # the next ordinary solve on the same wrapper starts from ITS graph, not from the diverged state
    ok = solve(make_graph([1.0, 1.0, 64, 64, 1e-3, 1.0, 1.0, 3200.0]))
    assert ok.newton_report.reason > 0


def test_in_place_edit_of_a_returned_graph_is_honoured():
    """A DeviceBackedGraph whose arrays were read can be edited in place (``phi[:] = ...``) without
    passing through a setter; the next solve must see the edit like it would on a plain Graph."""
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    solve = solver_builder.build('cuda', None)
    g1 = solve(make_graph([1.0, 1.0, 32, 32, 1e-2, 1.0, 1.0, 100.0]))
    a = solve(g1)                                   # resident path, untouched graph
    g1.ns_x_mesh.phi[:] *= 0.5                      # materialises, then edits in place
    b = solve(g1)
    plain = make_graph([1.0, 1.0, 32, 32, 1e-2, 1.0, 1.0, 100.0], None, g1.ns_x_mesh.phi, g1.ns_y_mesh.phi,
                       g1.pressure_mesh.phi)
    c = solve(plain)
    assert not np.allclose(a.ns_x_mesh.phi, b.ns_x_mesh.phi)
    assert np.array_equal(b.ns_x_mesh.phi, c.ns_x_mesh.phi)
    assert np.array_equal(b.ns_x_mesh.phi_old, g1.ns_x_mesh.phi)


def test_linear_solve_with_zero_rhs_returns_zero():
    import torch
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.lib()
    g = make_graph([1.0, 1.0, 32, 32, 1e-2, 1.0, 1.0, 100.0])
    grid = _native.grid_from_graph(g)
    h = ctypes.c_void_p()
    _native.check(lib.ldc_solver_create(ctypes.byref(grid), None, ctypes.byref(h)))
    try:
        st = _native.current_stream()
        _native.check(lib.ldc_solver_set_state(h, _native.ptr(torch.ones(3 * 32 * 32, dtype=torch.float64,
                                                                         device='cuda')), st))
        rhs = torch.zeros(3 * 32 * 32, dtype=torch.float64, device='cuda')
        y = torch.full_like(rhs, 7.0)
        its, reason, rel = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
        _native.check(lib.ldc_solver_linear_solve(h, _native.ptr(rhs), _native.ptr(y), ctypes.byref(its),
                                                  ctypes.byref(reason), ctypes.byref(rel), st))
        assert reason.value == 3 and its.value == 0
        assert float(y.abs().max()) == 0.0
    finally:
        lib.ldc_solver_destroy(h)


def test_time_ops_reports_positive_times():
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.lib()
    g = make_graph([1.0, 1.0, 128, 128, 1e-2, 1.0, 1.0, 100.0])
    grid = _native.grid_from_graph(g)
    h = ctypes.c_void_p()
    _native.check(lib.ldc_solver_create(ctypes.byref(grid), None, ctypes.byref(h)))
    try:
        import torch
        _native.check(lib.ldc_solver_set_state(h, _native.ptr(torch.ones(3 * 128 * 128, dtype=torch.float64,
                                                                         device='cuda')), _native.current_stream()))
        ms = (ctypes.c_double * 4)()
        _native.check(lib.ldc_solver_time_ops(h, 5, ms, _native.current_stream()))
        assert all(0.0 < v < 50.0 for v in ms), list(ms)
    finally:
        lib.ldc_solver_destroy(h)
