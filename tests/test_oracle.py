"""Pin the CPU oracle (oracle/) against the reference's own golden vectors.

Every fixture under tests/golden/ was produced by tests/golden/generate_golden.py from the
reference itself (Python package imported from /root/reference, C++ residuals compiled
unchanged into oracle/_ref, the reference's golden files).
"""
import os

import numpy as np
import pytest

from conftest import make_graph, synthetic_case
from oracle import ldc_oracle as orc


def test_reference_3x3_known_answer(golden):
    """tests/test_residual_function.py:8-44 of the reference; bit-exact."""
    d = golden('residual_3x3_reference_test.npz')
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    g = Graph(1.0, 1.0, 3, 3, 0.01, 1.0, 1.0, 100.0, initial_P=100.0, initial_U=0.001, initial_V=-0.01)
    F = orc.residual_function(d['X'], g)
    assert np.array_equal(F, d['F'])
    # the printed known-answer values of SURVEY App. E.1
    assert np.allclose(F[:3], [-6.0e-03, 1.50991111e-02, -1.51057778e-01], rtol=1e-8)
    assert np.allclose(F[19], -9.99848742e+01, rtol=1e-8)


def test_random_cases_bit_exact(golden):
    d = golden('residual_random_cases.npz')
    for k in range(int(d['count'])):
        g = make_graph(d['params_%d' % k], 0.9, d['U_%d' % k], d['V_%d' % k], d['P_%d' % k])
        for tag in ('', '_dummy'):
            X, F = d['X%s_%d' % (tag, k)], d['F%s_%d' % (tag, k)]
            assert np.array_equal(orc.residual_function(X, g, nthreads=1), F), (k, tag)
            assert np.array_equal(orc.residual_function(X, g, nthreads=0), F), (k, tag)
        assert np.array_equal(orc.create_X(g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi, g),
                              d['X_%d' % k])
        U, V, P = orc.recover_X(d['X_dummy_%d' % k], g)
        assert np.array_equal(U, d['U_%d' % k]) and np.array_equal(V, d['V_%d' % k])
        assert np.array_equal(P, d['P_%d' % k])


def test_mask_golden(golden):
    d = golden('jacobian_masks.npz')
    keys = sorted(k[len('indptr_'):] for k in d.files if k.startswith('indptr_'))
    assert len(keys) == 9
    for key in keys:
        nx, ny, dof = [int(v) for v in key.split('_')]
        indptr, indices = orc.jacobian_mask_arrays(nx, ny, dof)
        assert indptr.dtype == np.int32 and indices.dtype == np.int32
        assert np.array_equal(indptr, d['indptr_' + key]), key
        assert np.array_equal(indices, d['indices_' + key]), key
        if 'dense_' + key in d.files:
            shape = tuple(d['dense_shape_' + key])
            dense = np.unpackbits(d['dense_' + key])[:shape[0] * shape[1]].reshape(shape)
            assert np.array_equal(orc.jacobian_mask(nx, ny, dof).toarray(), dense)


def test_mask_nnz_closed_form():
    # nnz = 9(7N - 4nx) for dof=3
    for nx, ny in [(500, 500), (1024, 1024), (100, 100), (64, 64)]:
        from oracle.ldc_oracle import _lib
        assert _lib().ldc_oracle_mask_nnz(nx, ny, 3) == 9 * (7 * nx * ny - 4 * nx)
    assert _lib().ldc_oracle_mask_nnz(4096, 4096, 3) == 1056817152


@pytest.mark.skipif(not os.path.isdir(os.path.join(os.path.dirname(orc.__file__), '_ref')),
                    reason="oracle/_ref not built")
def test_against_compiled_reference():
    """oracle/_ref = reference C++ (serial and OpenMP) compiled unchanged."""
    ref = orc.ref_module()
    for nx, ny in [(100, 100), (256, 256), (257, 131)]:
        g, X = synthetic_case(nx, ny)
        F_ref = np.asarray(ref.residual_function(X, g))
        assert np.array_equal(np.asarray(ref.residual_function_omp(X, g)), F_ref)
        assert np.array_equal(orc.residual_function(X, g, nthreads=0), F_ref)
    g, X = synthetic_case(64, 48)
    g.use_cds = True   # only the numba backend of the reference honours it; C++ ignores it
    assert not np.array_equal(orc.residual_function(X, g), np.asarray(ref.residual_function(X, g)))


def test_central_differencing_against_reference_numba_backend(golden):
    """graph.use_cds: golden vectors from the reference's numba backend, the only one that honours
    the switch (numba_residual_function.py:128-137,202-211,247); bit-exact on and off"""
    d = golden('residual_cds_cases.npz')
    for k in range(int(d['count'])):
        g = make_graph(d['params_%d' % k])
        g.ns_x_mesh.phi_old = d['U_old_%d' % k]
        g.ns_y_mesh.phi_old = d['V_old_%d' % k]
        g.use_cds = False
        assert np.array_equal(orc.residual_function(d['X_%d' % k], g), d['F_uds_%d' % k]), k
        g.use_cds = True
        assert np.array_equal(orc.residual_function(d['X_%d' % k], g), d['F_cds_%d' % k]), k
        assert np.array_equal(orc.residual_function(d['X_%d' % k], g, nthreads=0), d['F_cds_%d' % k]), k


def test_fd_jacobian_colouring_equals_column_by_column():
    """coloured 'ds' FD + zeroed wrap entries == one residual per column."""
    for nx, ny in [(9, 8), (7, 7), (13, 9), (4, 3)]:
        g, X = synthetic_case(nx, ny, bc=100.0, seed=nx * 100 + ny)
        mask = orc.jacobian_mask_arrays(nx, ny, 3)
        F0 = orc.residual_function(X, g)
        a = orc.fd_jacobian(X, g, F0=F0, mask=mask)
        b = orc.fd_jacobian(X, g, F0=F0, mask=mask, by_column=True)
        assert np.array_equal(a, b), (nx, ny)
        # structure facts of SURVEY E.3: continuity rows never see P
        indptr, indices = mask
        for r in range(0, 3 * nx * ny, 3):
            cols = indices[indptr[r]:indptr[r + 1]]
            assert np.all(a[indptr[r]:indptr[r + 1]][cols % 3 == 0] == 0.0)


def test_newton_restatement_small_case(golden):
    """15x15 Re=10 end-to-end fixture of the reference (tests/test_run_solver.py:20-52),
    PETSc-equivalent restatement and the direct-solve Newton."""
    d = golden('small_case_15x15.npz')
    for linear in ('gmres_ilu0', 'direct'):
        g = make_graph([1.0, 1.0, 15, 15, 0.05, 1.0, 1.0, 10.0])
        res, hist = orc.run_simulation(g, 0.1, linear=linear)
        assert len(hist) == 2
        # the reference's own tolerance
        assert np.allclose(res.ns_x_mesh.phi, d['U_shipped'], rtol=1e-4, atol=1e-1)
        assert np.allclose(res.ns_y_mesh.phi, d['V_shipped'], rtol=1e-4, atol=1e-1)
    # tight Newton reproduces the (tightly converged) scipy fixture far below that
    g = make_graph([1.0, 1.0, 15, 15, 0.05, 1.0, 1.0, 10.0])
    res, _ = orc.run_simulation(g, 0.1, linear='direct', rtol=1e-13, atol=1e-11, stol=1e-14)
    assert np.abs(res.ns_x_mesh.phi - d['U_shipped']).max() < 1e-7
    assert np.abs(res.ns_y_mesh.phi - d['V_shipped']).max() < 1e-7


def test_ilu0_gmres_restatement_solves_the_linear_system():
    """self-consistency of the PETSc-equivalent restatement: ILU(0) (NONZERO shift) + left-
    preconditioned GMRES(30) solves J y = F to the requested tolerance (true residual checked
    with an independent CSR product), and needs the shift because continuity rows have a zero
    diagonal."""
    import scipy.sparse as sp
    g = make_graph([1.0, 1.0, 24, 20, 0.02, 1.0, 1.0, 100.0])
    X = orc.create_X(g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi, g)
    mask = orc.jacobian_mask_arrays(24, 20, 3)
    F = orc.residual_function(X, g)
    J = orc.fd_jacobian(X, g, F0=F, mask=mask)
    A = sp.csr_matrix((J, mask[1], mask[0]), shape=(X.size, X.size))
    assert np.all(A.diagonal()[0::3] == 0.0)              # zero pivots without the shift
    lu, diag, nshift, shift = orc.ilu0(mask, J)
    assert nshift >= 1 and shift > 0.0
    z = orc.ilu0_solve(mask, lu, diag, F)
    assert np.isfinite(z).all()
    y, reason, its, rnorm = orc.gmres_ilu0(mask, J, lu, diag, F, rtol=1e-10, max_it=2000)
    assert reason in (2, 3) and its > 0
    assert np.linalg.norm(F - A @ y) <= 1e-6 * np.linalg.norm(F)
    assert np.allclose(orc.spmv(mask, J, y), A @ y, rtol=0, atol=1e-9 * np.abs(A @ y).max())
