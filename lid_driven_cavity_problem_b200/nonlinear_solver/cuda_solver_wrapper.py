"""Device-resident Newton-Krylov solver plug-in -- the `cuda` sibling of
``PetscSolverWrapper`` / ``ScipySolverWrapper``
(nonlinear_solver/petsc_solver_wrapper.py:35-145, scipy_solver_wrapper.py:13-53).

``CudaSolverWrapper(residual_f).solve(graph) -> new Graph`` keeps the reference contract
(SURVEY 8b): the input graph is not modified, the returned graph has ``phi_old`` = the input's
``phi`` and ``phi`` = the converged fields, ``SolverDivergedException`` is raised on a diverged
Newton solve unless ``options.IGNORE_DIVERGED``, details are logged if
``options.SHOW_SOLVER_DETAILS``, and the wrapper object is reusable across time steps.

What differs from the PETSc wrapper is where things live: the state, Jacobian, Krylov basis and
multigrid hierarchy stay in HBM inside one ``ldc_solver`` handle that is created once (the
reference rebuilds mask, colouring and SNES every step).  The graph handed back is a
``DeviceBackedGraph`` whose meshes download their arrays only when somebody reads them; when the
time stepper passes that same graph back in, nothing is uploaded either.
"""
import ctypes
import logging

import numpy as np

from lid_driven_cavity_problem_b200 import _native, options
from lid_driven_cavity_problem_b200.nonlinear_solver.exceptions import SolverDivergedException
from lid_driven_cavity_problem_b200.staggered_grid import Graph

logger = logging.getLogger(__name__)

# SNESConvergedReason, as tabulated by the reference (petsc_solver_wrapper.py:12-31), plus -9
NONLINEAR_SOLVER_CONVERGENCE_REASONS = {
    0: 'still iterating',
    2: '||F|| < atol',
    3: '||F|| < rtol',
    4: 'Newton computed step size small; || delta x || < stol || x ||',
    5: 'maximum iterations reached',
    7: 'trust region delta',
    -1: 'the new x location passed to the function is not in the domain of F',
    -2: 'maximum function count reached',
    -3: 'the linear solver failed',
    -4: 'norm of F is NaN',
    -5: 'maximum iterations reached',
    -6: 'the line search failed',
    -7: 'inner solve failed',
    -8: '|| J^T b || is small, implies converged to local minimum of F()',
    -9: '||F|| > divtol * ||F_0||',
}


class _LazyMesh(object):
    """Mesh2d look-alike (nx, ny, phi, phi_old, len) whose arrays are fetched from the device
    snapshot of the owning graph on first access."""
    __slots__ = ('nx', 'ny', '_owner', '_field', '_phi', '_phi_old')

    def __init__(self, nx, ny, owner, field):
        self.nx, self.ny = nx, ny
        self._owner, self._field = owner, field
        self._phi = None
        self._phi_old = None

    def __len__(self):
        return self.nx * self.ny

    @property
    def phi(self):
        if self._phi is None:
            self._owner._materialise()
        return self._phi

    @phi.setter
    def phi(self, value):
        self._owner._materialise()
        self._phi = value
        self._owner._device_token = None   # host edit: the device copy is stale

    @property
    def phi_old(self):
        if self._phi_old is None:
            self._owner._materialise()
        return self._phi_old

    @phi_old.setter
    def phi_old(self, value):
        self._owner._materialise()
        self._phi_old = value


class DeviceBackedGraph(Graph):
    """A ``Graph`` returned by the cuda solver: same attributes as the reference's
    (staggered_grid.py:22-47), arrays materialised lazily from device snapshots."""
    __slots__ = ('_X_dev', '_X_old_dev', '_device_token', '_du_inf', 'newton_report', '_done')

    def __init__(self, template, X_dev, X_old_dev, token, du_inf, report):
        # no call to Graph.__init__: copy the scalars, build lazy meshes
        for name in ('dt', 'dx', 'dy', 'rho', 'mi', 'bc', 'use_cds'):
            setattr(self, name, getattr(template, name))
        nx, ny = template.pressure_mesh.nx, template.pressure_mesh.ny
        self.pressure_mesh = _LazyMesh(nx, ny, self, 'P')
        self.ns_x_mesh = _LazyMesh(nx - 1, ny, self, 'U')
        self.ns_y_mesh = _LazyMesh(nx, ny - 1, self, 'V')
        self._X_dev, self._X_old_dev = X_dev, X_old_dev
        self._device_token = token
        self._du_inf = du_inf
        self.newton_report = report
        self._done = False

    def _materialise(self):
        if self._done:
            return
        self._done = True
        # host arrays are about to be handed out and can be edited in place (``phi[:] = ...``,
        # ``phi *= 0.5``) without passing through a setter: from here on the device copy cannot be
        # trusted, so the next solve() uploads the host arrays like it does for a plain Graph
        self._device_token = None
        from lid_driven_cavity_problem_b200.nonlinear_solver._common import recover_X_device
        U, V, P = [t.cpu().numpy() for t in recover_X_device(self._X_dev, self)]
        Uo, Vo, Po = [t.cpu().numpy() for t in recover_X_device(self._X_old_dev, self)]
        for mesh, new, old in ((self.ns_x_mesh, U, Uo), (self.ns_y_mesh, V, Vo),
                               (self.pressure_mesh, P, Po)):
            if mesh._phi is None:
                mesh._phi = new
            if mesh._phi_old is None:
                mesh._phi_old = old

    def steady_state_norm(self):
        """max|U - U_old| / bc, computed on the device during the solve (time_stepper.py:39)."""
        return self._du_inf / self.bc


class CudaSolverWrapper(object):
    def __init__(self, residual_f=None, **solver_options):
        """``residual_f`` is accepted for plug-compatibility with ``solver_builder.build``; the
        device-resident solve always evaluates the residual with the CUDA kernel (there is no
        host callback inside the Newton loop).  ``solver_options`` override fields of
        ``ldc_solver_options`` (snes_rtol, ksp_rtol, pc_type, ...)."""
        self._residual_f = residual_f
        self._solver_options = dict(solver_options)
        self._handle = None
        self._key = None
        self._token = 0
        self._active_graph = None
        self.last_report = None
        self.total_linear_iterations = 0
        self.total_newton_iterations = 0
        self.total_function_evaluations = 0
        self.total_jacobian_evaluations = 0
        if residual_f is not None and getattr(residual_f, '__module__', '').split('.')[-1] != 'cuda_residual_function':
            logger.info("cuda solver: residual_f=%r is ignored, the device residual kernel is used", residual_f)

    # -- handle management -------------------------------------------------------------------
    def _options_struct(self):
        lib = _native.lib()
        opt = _native.LdcSolverOptions()
        lib.ldc_solver_default_options(ctypes.byref(opt))
        for k, v in self._solver_options.items():
            if not hasattr(opt, k):
                raise TypeError("unknown solver option %r" % k)
            setattr(opt, k, v)
        return opt

    def _ensure_handle(self, graph):
        nx, ny = int(graph.pressure_mesh.nx), int(graph.pressure_mesh.ny)
        key = (nx, ny, float(graph.dx), float(graph.dy), float(graph.rho), float(graph.mi),
               float(graph.bc), bool(getattr(graph, 'use_cds', False)))
        if self._handle is not None and key == self._key:
            return
        self.close()
        lib = _native.lib()
        g = _native.grid_from_graph(graph)
        opt = self._options_struct()
        handle = ctypes.c_void_p()
        _native.check(lib.ldc_solver_create(ctypes.byref(g), ctypes.byref(opt), ctypes.byref(handle)),
                      'ldc_solver_create')
        self._handle, self._key = handle, key
        self._token += 1
        logger.info("cuda solver: %dx%d grid, Jacobian NNZ=%d, %d multigrid level(s)", nx, ny,
                    lib.ldc_mask_nnz(nx, ny, 3), lib.ldc_solver_num_levels(handle))

    def close(self):
        if self._handle is not None:
            _native.lib().ldc_solver_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the plug point ----------------------------------------------------------------------
    def solve(self, graph):
        assert self._active_graph is None
        torch = _native.require_cuda()
        lib = _native.lib()
        self._ensure_handle(graph)
        self._active_graph = graph
        try:
            nx, ny = int(graph.pressure_mesh.nx), int(graph.pressure_mesh.ny)
            stream = _native.current_stream()
            resident = (isinstance(graph, DeviceBackedGraph) and graph._device_token is not None
                        and graph._device_token == (id(self), self._token))
            if not resident:
                from lid_driven_cavity_problem_b200.nonlinear_solver._common import create_X_device
                dev = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).cuda()
                       for a in (graph.ns_x_mesh.phi, graph.ns_y_mesh.phi, graph.pressure_mesh.phi)]
                X0 = create_X_device(dev[0], dev[1], dev[2], graph)
                _native.check(lib.ldc_solver_set_state(self._handle, _native.ptr(X0), stream), 'set_state')
            dt = (ctypes.c_double * 1)(float(graph.dt))
            _native.check(lib.ldc_solver_set_dt(self._handle, dt), 'set_dt')

            X_old = torch.empty(3 * nx * ny, dtype=torch.float64, device='cuda')
            _native.check(lib.ldc_solver_get_state(self._handle, _native.ptr(X_old), stream), 'get_state')
            # options.IGNORE_DIVERGED (read at call time like the reference does): a diverged solve
            # hands back its last iterate (petsc_solver_wrapper.py:90-103), so the library must not
            # roll the state back
            _native.check(lib.ldc_solver_set_keep_diverged(self._handle, 1 if options.IGNORE_DIVERGED else 0),
                          'ldc_solver_set_keep_diverged')
            report = _native.LdcNewtonReport()
            du = (ctypes.c_double * 1)(0.0)
            _native.check(lib.ldc_solver_step(self._handle, None, ctypes.byref(report), du, stream),
                          'ldc_solver_step')
            self.last_report = report
            self.total_linear_iterations += report.linear_iterations
            self.total_newton_iterations += report.iterations
            self.total_function_evaluations += report.function_evaluations
            self.total_jacobian_evaluations += report.jacobian_evaluations

            if options.SHOW_SOLVER_DETAILS:
                logger.info("Number of function calls=%s, FD Jacobian launches=%s (Newton its=%s, Krylov its=%s)",
                            report.function_evaluations, report.jacobian_evaluations, report.iterations,
                            report.linear_iterations)
                text = NONLINEAR_SOLVER_CONVERGENCE_REASONS.get(report.reason, '?')
                if report.reason > 0:
                    logger.info("Converged with reason: %s", text)
                else:
                    logger.info("Diverged with reason: %s", text)
                    if report.reason == -3:
                        logger.info("Linear solver divergence reason code: %s", report.ksp_reason)

            if report.reason <= 0 and not options.IGNORE_DIVERGED:
                # the device state has been rolled back to the entry state by the library
                self._token += 1
                raise SolverDivergedException()

            X_new = torch.empty_like(X_old)
            _native.check(lib.ldc_solver_get_state(self._handle, _native.ptr(X_new), stream), 'get_state')
            self._token += 1
            return DeviceBackedGraph(graph, X_new, X_old, (id(self), self._token), du[0], report)
        finally:
            self._active_graph = None
