"""CPU ORACLE -- test infrastructure, NOT the product.

ctypes front-end of ``oracle/libldc_oracle.so`` (the C restatement in ldc_oracle.c) plus
the PETSc-equivalent Newton / time-step restatement in numpy.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product package never does.

Parity status: residual / packing / mask are pinned against the reference itself
(tests/test_oracle.py, tests/golden/).  FD Jacobian, ILU(0), GMRES and the Newton loop
restate PETSc behaviour selected by nonlinear_solver/petsc_solver_wrapper.py:112-145 and
are PARITY UNPINNED (PETSc/petsc4py cannot be installed in this image).
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libldc_oracle.so')
        if not os.path.exists(path):
            raise RuntimeError("oracle library is not compiled (run `make -C oracle oracle` "
                               "or `python -c 'import __graft_entry__ as g; g.build()'`)")
        lib = ctypes.CDLL(path)
        lib.ldc_oracle_residual.restype = ctypes.c_int
        lib.ldc_oracle_residual.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_long, ctypes.c_long] + \
            [ctypes.c_double] * 6 + [ctypes.c_int, ctypes.c_int]
        lib.ldc_oracle_max_threads.restype = ctypes.c_int
        lib.ldc_oracle_create_X.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_long, ctypes.c_long]
        lib.ldc_oracle_recover_X.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_long, ctypes.c_long]
        lib.ldc_oracle_mask_nnz.restype = ctypes.c_long
        lib.ldc_oracle_mask_nnz.argtypes = [ctypes.c_long] * 3
        lib.ldc_oracle_mask_csr.restype = ctypes.c_int
        lib.ldc_oracle_mask_csr.argtypes = [ctypes.c_long] * 3 + [_ip, _ip]
        fd_args = [_dp, _dp, _dp, _dp, _ip, _ip, _dp, ctypes.c_long, ctypes.c_long] + \
            [ctypes.c_double] * 6 + [ctypes.c_int]
        lib.ldc_oracle_fd_jacobian.restype = ctypes.c_int
        lib.ldc_oracle_fd_jacobian.argtypes = fd_args + [ctypes.c_int]
        lib.ldc_oracle_fd_jacobian_bycolumn.restype = ctypes.c_int
        lib.ldc_oracle_fd_jacobian_bycolumn.argtypes = fd_args
        lib.ldc_oracle_spmv.argtypes = [_ip, _ip, _dp, _dp, _dp, ctypes.c_long, ctypes.c_int]
        lib.ldc_oracle_ilu0.restype = ctypes.c_int
        lib.ldc_oracle_ilu0.argtypes = [_ip, _ip, _dp, _dp, _ip, ctypes.c_long, _dp]
        lib.ldc_oracle_ilu0_solve.argtypes = [_ip, _ip, _dp, _ip, _dp, _dp, ctypes.c_long]
        lib.ldc_oracle_gmres_ilu0.restype = ctypes.c_int
        lib.ldc_oracle_gmres_ilu0.argtypes = [_ip, _ip, _dp, _dp, _ip, _dp, _dp, ctypes.c_long,
                                              ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                              ctypes.c_double, ctypes.c_long,
                                              ctypes.POINTER(ctypes.c_long), _dp, ctypes.c_int]
        _LIB = lib
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return int(_lib().ldc_oracle_max_threads())


def _graph_scalars(graph):
    return (float(graph.dt), float(graph.dx), float(graph.dy), float(graph.rho), float(graph.mi),
            float(graph.bc))


def residual_function(X, graph, nthreads=1):
    """F(X; graph) -- same contract as the reference's residual_function modules
    (c++/residual_function.cpp:10).  ``nthreads=0`` uses every OpenMP thread."""
    X = _f64(X)
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    assert X.size == 3 * nx * ny
    U_old = _f64(graph.ns_x_mesh.phi_old)
    V_old = _f64(graph.ns_y_mesh.phi_old)
    F = np.empty_like(X)
    rc = _lib().ldc_oracle_residual(_d(X), _d(U_old), _d(V_old), _d(F), nx, ny,
                                    *_graph_scalars(graph),
                                    int(bool(getattr(graph, 'use_cds', False))), int(nthreads))
    if rc != 0:
        raise ValueError("oracle residual rejected the grid %sx%s" % (nx, ny))
    return F


def create_X(U, V, P, graph):
    """nonlinear_solver/_common.py:4-18"""
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    U, V, P = _f64(U), _f64(V), _f64(P)
    X = np.empty(3 * nx * ny)
    _lib().ldc_oracle_create_X(_d(U), _d(V), _d(P), _d(X), nx, ny)
    return X


def recover_X(X, graph):
    """nonlinear_solver/_common.py:21-36; returns (U, V, P)"""
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    X = _f64(X)
    U = np.empty((nx - 1) * ny)
    V = np.empty(nx * (ny - 1))
    P = np.empty(nx * ny)
    _lib().ldc_oracle_recover_X(_d(X), _d(U), _d(V), _d(P), nx, ny)
    return U, V, P


def jacobian_mask_arrays(nx, ny, dof=3):
    """(indptr, indices) of _calculate_jacobian_mask (nonlinear_solver/_common.py:39-57)."""
    nnz = _lib().ldc_oracle_mask_nnz(nx, ny, dof)
    indptr = np.empty(dof * nx * ny + 1, dtype=np.int32)
    indices = np.empty(nnz, dtype=np.int32)
    rc = _lib().ldc_oracle_mask_csr(nx, ny, dof, _i(indptr), _i(indices))
    if rc != 0:
        raise ValueError("mask needs nx >= 3")
    return indptr, indices


def jacobian_mask(nx, ny, dof=3):
    from scipy.sparse import csr_matrix
    indptr, indices = jacobian_mask_arrays(nx, ny, dof)
    n = dof * nx * ny
    return csr_matrix((np.ones(indices.size), indices, indptr), shape=(n, n))


def fd_jacobian(X, graph, F0=None, mask=None, nthreads=0, by_column=False):
    """CSR values of the 'ds' finite-difference Jacobian on the mask pattern."""
    X = _f64(X)
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    if mask is None:
        mask = jacobian_mask_arrays(nx, ny, 3)
    indptr, indices = mask
    if F0 is None:
        F0 = residual_function(X, graph)
    F0 = _f64(F0)
    U_old = _f64(graph.ns_x_mesh.phi_old)
    V_old = _f64(graph.ns_y_mesh.phi_old)
    values = np.zeros(indices.size)
    args = [_d(X), _d(U_old), _d(V_old), _d(F0), _i(indptr), _i(indices), _d(values), nx, ny,
            *_graph_scalars(graph), int(bool(getattr(graph, 'use_cds', False)))]
    if by_column:
        rc = _lib().ldc_oracle_fd_jacobian_bycolumn(*args)
    else:
        rc = _lib().ldc_oracle_fd_jacobian(*args, int(nthreads))
    assert rc == 0
    return values


def spmv(mask, values, x, nthreads=0):
    indptr, indices = mask
    x = _f64(x)
    y = np.empty_like(x)
    _lib().ldc_oracle_spmv(_i(indptr), _i(indices), _d(values), _d(x), _d(y), x.size, nthreads)
    return y


def ilu0(mask, values):
    indptr, indices = mask
    n = indptr.size - 1
    lu = np.empty_like(values)
    diag = np.empty(n, dtype=np.int32)
    shift = ctypes.c_double(0.0)
    rc = _lib().ldc_oracle_ilu0(_i(indptr), _i(indices), _d(values), _d(lu), _i(diag), n,
                                ctypes.byref(shift))
    if rc < 0:
        raise RuntimeError("ILU(0) failed with %d" % rc)
    return lu, diag, rc, shift.value


def ilu0_solve(mask, lu, diag, r):
    indptr, indices = mask
    r = _f64(r)
    z = np.empty_like(r)
    _lib().ldc_oracle_ilu0_solve(_i(indptr), _i(indices), _d(lu), _i(diag), _d(r), _d(z), r.size)
    return z


def gmres_ilu0(mask, values, lu, diag, b, restart=30, rtol=1e-5, atol=1e-50, dtol=1e5,
               max_it=10000, nthreads=0):
    """Left-preconditioned GMRES(restart)+ILU(0), PETSc KSP defaults.
    Returns (x, reason, iterations, preconditioned residual norm)."""
    indptr, indices = mask
    b = _f64(b)
    x = np.zeros_like(b)
    its = ctypes.c_long(0)
    rn = ctypes.c_double(0.0)
    reason = _lib().ldc_oracle_gmres_ilu0(_i(indptr), _i(indices), _d(values), _d(lu), _i(diag),
                                          _d(b), _d(x), b.size, restart, rtol, atol, dtol, max_it,
                                          ctypes.byref(its), ctypes.byref(rn), nthreads)
    return x, reason, its.value, rn.value


# --------------------------------------------------------------------------------------
# Newton (SNES newtonls, line search 'basic') -- petsc_solver_wrapper.py:112-145.
# PARITY UNPINNED: restates PETSc's SNESSolve_NEWTONLS + SNESConvergedDefault.
# --------------------------------------------------------------------------------------
class NewtonStats(object):
    def __init__(self):
        self.reason = 0
        self.iterations = 0
        self.function_evaluations = 0
        self.linear_iterations = 0
        self.ksp_reason = 0
        self.fnorms = []


def newton_solve(X0, graph, rtol=1e-4, atol=1e-4, stol=1e-4, max_it=50, max_funcs=10000,
                 linear='gmres_ilu0', ksp_rtol=1e-5, ksp_max_it=10000, nthreads=0, mask=None,
                 residual_f=None):
    """Full-step Newton on F(X)=0.  ``linear`` is 'gmres_ilu0' (the PETSc configuration of the
    reference) or 'direct' (scipy sparse LU with the pressure level pinned -- an independent
    tightly-converging solve used for converged-state parity).  Returns (X, NewtonStats)."""
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    if mask is None:
        mask = jacobian_mask_arrays(nx, ny, 3)
    if residual_f is None:
        def residual_f(x, g):
            return residual_function(x, g, nthreads=nthreads)
    st = NewtonStats()
    x = _f64(X0).copy()
    F = np.asarray(residual_f(x, graph), dtype=np.float64)
    st.function_evaluations += 1
    fnorm = float(np.linalg.norm(F))
    st.fnorms.append(fnorm)
    if fnorm != fnorm:
        st.reason = -4
        return x, st
    if fnorm < atol:
        st.reason = 2
        return x, st
    ttol = fnorm * rtol
    for it in range(max_it):
        J = fd_jacobian(x, graph, F0=F, mask=mask, nthreads=nthreads)
        st.function_evaluations += 21
        if linear == 'direct':
            y = _direct_solve(mask, J, F)
        else:
            lu, diag, _, _ = ilu0(mask, J)
            y, kreason, kits, _ = gmres_ilu0(mask, J, lu, diag, F, rtol=ksp_rtol,
                                             max_it=ksp_max_it, nthreads=nthreads)
            st.linear_iterations += kits
            st.ksp_reason = kreason
            if kreason < 0:
                st.reason = -3
                st.iterations = it
                return x, st
        x = x - y
        F = np.asarray(residual_f(x, graph), dtype=np.float64)
        st.function_evaluations += 1
        fnorm = float(np.linalg.norm(F))
        st.fnorms.append(fnorm)
        st.iterations = it + 1
        ynorm = float(np.linalg.norm(y))
        xnorm = float(np.linalg.norm(x))
        if fnorm != fnorm:
            st.reason = -4
            return x, st
        if fnorm < atol:
            st.reason = 2
            return x, st
        if st.function_evaluations >= max_funcs:
            st.reason = -2
            return x, st
        if fnorm <= ttol:
            st.reason = 3
            return x, st
        if ynorm < stol * xnorm:
            st.reason = 4
            return x, st
    st.reason = -5
    return x, st


def _direct_solve(mask, values, rhs):
    """J y = rhs by sparse LU.  J is singular by the pressure constant: the
    continuity row of cell 0 is replaced by 'P[0] correction = 0'."""
    from scipy.sparse import csr_matrix, lil_matrix
    from scipy.sparse.linalg import splu
    indptr, indices = mask
    n = indptr.size - 1
    A = csr_matrix((values, indices, indptr), shape=(n, n)).tolil()
    A[0, :] = 0.0
    A[0, 0] = 1.0
    b = np.array(rhs, dtype=np.float64)
    b[0] = 0.0
    return splu(A.tocsc()).solve(b)


class SolverDiverged(RuntimeError):
    pass


def solve_step(graph, **newton_kwargs):
    """One pseudo-time step; mirrors PetscSolverWrapper.solve (petsc_solver_wrapper.py:50-103):
    phi_old <- phi, Newton from X=phi, returns (new graph, NewtonStats)."""
    from copy import deepcopy
    g = deepcopy(graph)
    g.ns_x_mesh.phi_old = np.array(graph.ns_x_mesh.phi, dtype=np.float64)
    g.ns_y_mesh.phi_old = np.array(graph.ns_y_mesh.phi, dtype=np.float64)
    g.pressure_mesh.phi_old = np.array(graph.pressure_mesh.phi, dtype=np.float64)
    X0 = create_X(g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi, g)
    X, st = newton_solve(X0, g, **newton_kwargs)
    if st.reason <= 0:
        raise SolverDiverged(st.reason)
    U, V, P = recover_X(X, g)
    g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi = U, V, P
    return g, st


def run_simulation(graph, final_time, minimum_dt=1e-6, adaptative_dt=True, max_steps=None,
                   on_step=None, **newton_kwargs):
    """time_stepper.py:9-56 driving :func:`solve_step`.  Returns (graph, list of NewtonStats)."""
    t = 0.0
    history = []
    while final_time is None or t < final_time:
        if final_time is not None and t + graph.dt > final_time:
            graph.dt = final_time - t
        try:
            new_graph, st = solve_step(graph, **newton_kwargs)
        except SolverDiverged:
            if adaptative_dt:
                graph.dt /= 2.0
            if graph.dt <= minimum_dt:
                raise RuntimeError("Timestep has reached a too low value. Giving up.")
            continue
        graph = new_graph
        t += graph.dt
        history.append(st)
        if on_step is not None:
            on_step(graph, t, st)
        if final_time is None:
            norm = np.linalg.norm(graph.ns_x_mesh.phi - graph.ns_x_mesh.phi_old, ord=np.inf) / graph.bc
            if norm < 1e-6:
                break
            if adaptative_dt:
                graph.dt *= 1.2
        elif abs(final_time - t) <= 1e-5:
            if adaptative_dt:
                graph.dt *= 2.0
        if max_steps is not None and len(history) >= max_steps:
            break
    return graph, history


# --------------------------------------------------------------------------------------
# oracle/_ref : the reference's own C++ residuals, compiled unchanged (oracle/Makefile).
# --------------------------------------------------------------------------------------
def ref_module():
    """Import ``_residual_function`` built from /root/reference/c++ (serial + OpenMP)."""
    ref_dir = os.path.join(_HERE, '_ref')
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import _residual_function
    except ImportError as e:
        raise RuntimeError("oracle/_ref is not built (needs /root/reference at build time): %s" % e)
    return _residual_function
