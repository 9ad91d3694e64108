"""Packing between the three meshes and the interleaved unknown vector, and the Jacobian
sparsity mask -- the cuda backend's versions of the reference helpers
``_create_X`` / ``_recover_X`` / ``_calculate_jacobian_mask`` (nonlinear_solver/_common.py).

Host (numpy) callers get numpy / scipy objects like in the reference; the work is done by the
device kernels ``ldc_pack_X`` / ``ldc_unpack_X`` / ``ldc_mask_csr``.
"""
import numpy as np

from lid_driven_cavity_problem_b200 import _native


def _mesh_sizes(graph):
    nx, ny = int(graph.pressure_mesh.nx), int(graph.pressure_mesh.ny)
    return nx, ny, nx * ny, (nx - 1) * ny, nx * (ny - 1)


def create_X_device(U_dev, V_dev, P_dev, graph, batch=1):
    torch = _native.require_cuda()
    nx, ny, N, nU, nV = _mesh_sizes(graph)
    X = torch.empty(3 * N * batch, dtype=torch.float64, device='cuda')
    g = _native.grid_from_graph(graph, batch=batch)
    _native.check(_native.lib().ldc_pack_X(g, _native.ptr(U_dev), _native.ptr(V_dev), _native.ptr(P_dev),
                                           _native.ptr(X), _native.current_stream()), 'ldc_pack_X')
    return X


def recover_X_device(X_dev, graph, batch=1):
    torch = _native.require_cuda()
    nx, ny, N, nU, nV = _mesh_sizes(graph)
    U = torch.empty(nU * batch, dtype=torch.float64, device='cuda')
    V = torch.empty(nV * batch, dtype=torch.float64, device='cuda')
    P = torch.empty(N * batch, dtype=torch.float64, device='cuda')
    g = _native.grid_from_graph(graph, batch=batch)
    _native.check(_native.lib().ldc_unpack_X(g, _native.ptr(X_dev), _native.ptr(U), _native.ptr(V),
                                             _native.ptr(P), _native.current_stream()), 'ldc_unpack_X')
    return U, V, P


def _create_X(U, V, P, graph):
    """numpy in, numpy out (reference: _common.py:4-18)."""
    torch = _native.require_cuda()
    nx, ny, N, nU, nV = _mesh_sizes(graph)
    dev = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).cuda()
           for a in (U, V, P)]
    if dev[0].numel() != nU or dev[1].numel() != nV or dev[2].numel() != N:
        raise ValueError("mesh arrays do not match the %dx%d grid" % (nx, ny))
    return create_X_device(dev[0], dev[1], dev[2], graph).cpu().numpy()


def _recover_X(X, graph):
    """numpy in, (U, V, P) numpy out (reference: _common.py:21-36)."""
    torch = _native.require_cuda()
    nx, ny, N, nU, nV = _mesh_sizes(graph)
    X_dev = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64).reshape(-1)).cuda()
    if X_dev.numel() != 3 * N:
        raise ValueError("X has %d entries, expected %d" % (X_dev.numel(), 3 * N))
    return tuple(t.cpu().numpy() for t in recover_X_device(X_dev, graph))


def jacobian_mask_device(nx, ny, dof=3):
    """(indptr, indices) int32 CUDA tensors of the mask's canonical CSR."""
    torch = _native.require_cuda()
    lib = _native.lib()
    nnz = int(lib.ldc_mask_nnz(nx, ny, dof))
    if nnz < 0:
        raise ValueError("mask needs nx >= 3 (offsets collide otherwise)")
    indptr = torch.empty(dof * nx * ny + 1, dtype=torch.int32, device='cuda')
    indices = torch.empty(nnz, dtype=torch.int32, device='cuda')
    _native.check(lib.ldc_mask_csr(nx, ny, dof, _native.ptr(indptr), _native.ptr(indices),
                                   _native.current_stream()), 'ldc_mask_csr')
    return indptr, indices


def _calculate_jacobian_mask(nx, ny, dof):
    """scipy CSR with data == 1.0, int32 sorted indices (reference: _common.py:39-57)."""
    from scipy.sparse import csr_matrix
    indptr, indices = jacobian_mask_device(nx, ny, dof)
    indices = indices.cpu().numpy()
    n = dof * nx * ny
    return csr_matrix((np.ones(indices.size), indices, indptr.cpu().numpy()), shape=(n, n))
