"""Small-case solver plug-in on ``scipy.optimize.fsolve`` (dense MINPACK hybrd), kept beside
the cuda solver like in the reference (nonlinear_solver/scipy_solver_wrapper.py:13-53).  It is
the host-residual compatibility path: any ``residual_f(X, graph)`` works, including the cuda
residual with numpy in / numpy out, which makes the reference's 15x15 end-to-end test an
end-to-end check of the CUDA kernel alone.  Dense n x n storage limits it to <~ 60x60 grids."""
from copy import deepcopy
import logging

import numpy as np

from lid_driven_cavity_problem_b200 import options
from lid_driven_cavity_problem_b200.nonlinear_solver.exceptions import SolverDivergedException

logger = logging.getLogger(__name__)


def _pack(graph):
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    X = np.zeros(3 * nx * ny)
    X[0::3] = graph.pressure_mesh.phi
    Ue = np.zeros((ny, nx))
    Ue[:, :nx - 1] = np.asarray(graph.ns_x_mesh.phi).reshape(ny, nx - 1)
    X[1::3] = Ue.ravel()
    X[2:3 * nx * (ny - 1):3] = graph.ns_y_mesh.phi
    return X


def _unpack(X, graph):
    nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
    P = np.array(X[0::3])
    U = np.array(X[1::3].reshape(ny, nx)[:, :nx - 1]).ravel()
    V = np.array(X[2::3][:nx * (ny - 1)])
    return U, V, P


class ScipySolverWrapper(object):
    def __init__(self, residual_f):
        self._residual_f = residual_f

    def solve(self, graph):
        from scipy.optimize import fsolve
        g = deepcopy(graph)
        for mesh in (g.ns_x_mesh, g.ns_y_mesh, g.pressure_mesh):
            mesh.phi_old = np.array(mesh.phi, dtype=np.float64)
        X0 = _pack(g)
        X, info, ier, mesg = fsolve(self._residual_f, X0, args=(g,), full_output=True)
        if options.SHOW_SOLVER_DETAILS:
            logger.info("Number of function calls=%s", info['nfev'])
            logger.info("Converged" if ier == 1 else "Diverged: %s" % mesg)
        if ier != 1 and not options.IGNORE_DIVERGED:
            raise SolverDivergedException()
        g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi = _unpack(X, g)
        return g
