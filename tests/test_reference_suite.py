"""The reference's own three test files, case for case, with the `cuda` backend plugged in
(reference: tests/test_residual_function.py, tests/test_calculate_jacobian_mask.py,
tests/test_run_solver.py).  Golden data: tests/golden/ (generated from the reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_residual_function(golden):
    """tests/test_residual_function.py:8-44 with cuda_residual_function added to functions_to_try"""
    from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    x = np.array([123.0, 0.002, -0.02] * 9)
    graph = Graph(1.0, 1.0, 3, 3, 0.01, 1.0, 1.0, 100.0, initial_P=100.0, initial_U=0.001, initial_V=-0.01)
    functions_to_try = [cuda_residual_function.residual_function]
    reference_results = golden('residual_3x3_reference_test.npz')['F']   # pure-Python reference output
    for f in functions_to_try:
        assert np.allclose(f(x, graph), reference_results, rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize(("nx, ny, dof"), ((5, 5, 1), (4, 4, 3), (5, 4, 3), (4, 5, 3)))
def test_calculate_jacobian_mask(nx, ny, dof, golden):
    """tests/test_calculate_jacobian_mask.py:12-48 against the reference's golden J_*.txt"""
    from lid_driven_cavity_problem_b200.nonlinear_solver._common import _calculate_jacobian_mask
    d = golden('jacobian_masks.npz')
    key = '%d_%d_%d' % (nx, ny, dof)
    shape = tuple(d['dense_shape_' + key])
    expected_mask = np.unpackbits(d['dense_' + key])[:shape[0] * shape[1]].reshape(shape)
    mask = _calculate_jacobian_mask(nx, ny, dof).toarray()
    assert np.allclose(mask, expected_mask)


def test_smoke_medium_size_jacobian():
    from lid_driven_cavity_problem_b200.nonlinear_solver._common import _calculate_jacobian_mask
    assert _calculate_jacobian_mask(500, 500, 3).nnz == 15732000


@pytest.mark.parametrize(("solver_name"), ('cuda', 'scipy'))
def test_small_case(solver_name, golden):
    """tests/test_run_solver.py:13-52 (the reference parametrises petsc / scipy with the C++ residual)"""
    from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
    from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    size_x = size_y = 1.0
    nx = ny = 15
    dt, rho, final_time, mi, Re = 0.05, 1.0, 0.1, 1.0, 10
    U_bc = (mi * Re) / (rho * size_x)
    solver = solver_builder.build(solver_name, cuda_residual_function.residual_function)
    graph = Graph(size_x, size_y, nx, ny, dt, rho, mi, U_bc)
    result = run_simulation(graph, final_time, solver)
    U = np.array(result.ns_x_mesh.phi)
    V = np.array(result.ns_y_mesh.phi)
    d = golden('small_case_15x15.npz')
    assert np.allclose(U, d['U_shipped'], rtol=1e-4, atol=1e-1)
    assert np.allclose(V, d['V_shipped'], rtol=1e-4, atol=1e-1)


def test_jacobian_pattern_dump():
    """debug dump replacing _plot_jacobian: true pattern = 4 / 11 / 11 entries per interior cell
    (SURVEY E.3), a subset of the mask"""
    from conftest import synthetic_case
    from lid_driven_cavity_problem_b200.nonlinear_solver import _utils, _common
    from oracle import ldc_oracle as orc
    nx, ny = 12, 9
    g, X = synthetic_case(nx, ny, bc=100.0)
    P = _utils.jacobian_pattern(g, X)
    mask = _common._calculate_jacobian_mask(nx, ny, 3)
    assert P.nnz < mask.nnz and (P.astype(np.float64) - P.astype(np.float64).multiply(mask)).nnz == 0
    c = (nx // 2) + nx * (ny // 2)
    counts = np.diff(P.tocsr().indptr)
    assert list(counts[3 * c:3 * c + 3]) == [4, 11, 11]
    ref = orc.fd_jacobian(X, g)
    assert np.array_equal(_utils.jacobian_csr(g, X).data, ref)
