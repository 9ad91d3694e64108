"""Debug helper: the sparsity pattern of the Jacobian as the device sees it.

The reference's ``_plot_jacobian`` (nonlinear_solver/_utils.py:7-26) builds a dense forward-
difference Jacobian with a private scipy helper (gone from current scipy), thresholds it and shows
the boolean image, then the solver wrapper aborts (petsc_solver_wrapper.py:66-69).  Here the
coloured-FD Jacobian is assembled on the device into the reference mask's CSR and returned as a
scipy matrix (values) / boolean pattern -- no plotting dependency, usable on any grid size.
"""
import numpy as np

from lid_driven_cavity_problem_b200 import _native
from lid_driven_cavity_problem_b200.nonlinear_solver import _common


def jacobian_csr(graph, X=None):
    """scipy CSR (reference mask pattern, explicit zeros kept) of dF/dX at ``X`` (default: the
    graph's own fields), finite-differenced with PETSc's 'ds' rule on the device."""
    from scipy.sparse import csr_matrix
    torch = _native.require_cuda()
    lib = _native.lib()
    nx, ny = int(graph.pressure_mesh.nx), int(graph.pressure_mesh.ny)
    if X is None:
        X = _common._create_X(graph.ns_x_mesh.phi, graph.ns_y_mesh.phi, graph.pressure_mesh.phi, graph)
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).cuda()
    Ud = torch.from_numpy(np.ascontiguousarray(graph.ns_x_mesh.phi_old, dtype=np.float64)).cuda()
    Vd = torch.from_numpy(np.ascontiguousarray(graph.ns_y_mesh.phi_old, dtype=np.float64)).cuda()
    indptr, indices = _common.jacobian_mask_device(nx, ny, 3)
    values = torch.empty(indices.numel(), dtype=torch.float64, device='cuda')
    g = _native.grid_from_graph(graph)
    _native.check(lib.ldc_fd_jacobian(g, _native.ptr(Xd), _native.ptr(Ud), _native.ptr(Vd), _native.ptr(values),
                                      _native.current_stream()), 'ldc_fd_jacobian')
    n = 3 * nx * ny
    return csr_matrix((values.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(n, n))


def jacobian_pattern(graph, X=None, threshold=1e-15):
    """boolean CSR: entries with |J| > threshold (what the reference's image shows)"""
    J = jacobian_csr(graph, X)
    J.data = (np.abs(J.data) > threshold).astype(np.float64)
    J.eliminate_zeros()
    return J.astype(bool)
