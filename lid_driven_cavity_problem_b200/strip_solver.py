"""Row-strip domain decomposition of ONE large cavity over the GPUs of a node (BASELINE.json
config 5: 4096x4096 over 1/2/4/8 B200).  One process per GPU; rank r owns the grid rows
``strip_partition(ny, world, r)``; cells are row-major, so a strip is a contiguous slice of the
interleaved unknown vector and every stencil reaches at most one row across a strip boundary
(SURVEY 8e).  The device solver of each rank exchanges one halo row with its neighbours per
operator application and all-reduces norms / Gram-Schmidt coefficients (libldc_b200_comm.so,
NCCL over NVLink); the preconditioner is block-Jacobi over strips.  The reference has no
counterpart (it is a single sequential process); the time-stepper rules are those of
time_stepper.py:9-56, applied identically on every rank from all-reduced scalars.
"""
import ctypes
import logging

import numpy as np

from lid_driven_cavity_problem_b200 import _comm, _native
from lid_driven_cavity_problem_b200.nonlinear_solver.exceptions import SolverDivergedException

logger = logging.getLogger(__name__)


class CudaStripSolver(object):
    def __init__(self, nx, ny, dt, rho, mi, bc, size_x=1.0, size_y=1.0, comm=None, rank=0, world=1,
                 **solver_options):
        if ny % world != 0:
            # unequal strips would build different multigrid hierarchies per rank (different
            # sequences of collectives -> a hang); every rank sees the same ny / world, so every
            # rank refuses.  ldc_solver_set_comm checks the same thing collectively.
            raise ValueError("row strips must be equal: ny=%d is not divisible by %d ranks" % (ny, world))
        torch = _native.require_cuda()
        lib = _native.lib()
        self.nx, self.ny, self.bc, self.dt = int(nx), int(ny), float(bc), float(dt)
        self.rank, self.world, self.comm = rank, world, comm
        self.row_begin, self.row_count = _comm.strip_partition(ny, world, rank, self._align(ny, world))
        g = _native.LdcGrid()
        g.nx, g.ny, g.row_begin, g.row_count, g.batch, g.use_cds = nx, ny, self.row_begin, self.row_count, 1, 0
        g.dt, g.dx, g.dy, g.rho, g.mi, g.bc = float(dt), size_x / nx, size_y / ny, float(rho), float(mi), float(bc)
        opt = _native.LdcSolverOptions()
        lib.ldc_solver_default_options(ctypes.byref(opt))
        for k, v in solver_options.items():
            if not hasattr(opt, k):
                raise TypeError("unknown solver option %r" % k)
            setattr(opt, k, v)
        self._handle = ctypes.c_void_p()
        _native.check(lib.ldc_solver_create(ctypes.byref(g), ctypes.byref(opt), ctypes.byref(self._handle)),
                      'ldc_solver_create')
        if comm is not None:
            _native.check(lib.ldc_solver_set_comm(self._handle, comm.ops()), 'ldc_solver_set_comm')
        self.n_local = 3 * nx * self.row_count
        self.newton_iterations = self.krylov_iterations = self.steps = self.retries = 0
        self._torch = torch

    def close(self):
        if self._handle is not None:
            _native.lib().ldc_solver_destroy(self._handle)
            self._handle = None

    def set_uniform_state(self, P=1.0, U=1.0, V=1.0):
        """the reference's default start (staggered_grid.py:36): every real unknown = 1"""
        nx, rows = self.nx, self.row_count
        X = np.zeros((rows, nx, 3))
        X[:, :, 0] = P
        X[:, :nx - 1, 1] = U
        jg = self.row_begin + np.arange(rows)
        X[jg < self.ny - 1, :, 2] = V
        self.set_state(self._torch.from_numpy(X.reshape(-1)).cuda())

    def set_state(self, X_local_dev):
        assert X_local_dev.numel() == self.n_local
        _native.check(_native.lib().ldc_solver_set_state(self._handle, _native.ptr(X_local_dev),
                                                         _native.current_stream()), 'set_state')

    def get_state(self):
        X = self._torch.empty(self.n_local, dtype=self._torch.float64, device='cuda')
        _native.check(_native.lib().ldc_solver_get_state(self._handle, _native.ptr(X), _native.current_stream()),
                      'get_state')
        return X

    def gather_state(self):
        """full interleaved X on every rank (for checks / output of moderate grids)"""
        X = self.get_state()
        if self.world == 1:
            return X
        import torch.distributed as dist
        sizes = [3 * self.nx * _comm.strip_partition(self.ny, self.world, r, self._align(self.ny, self.world))[1]
                 for r in range(self.world)]
        parts = [self._torch.empty(sz, dtype=X.dtype, device='cuda') for sz in sizes]
        dist.all_gather(parts, X)
        return self._torch.cat(parts)

    @staticmethod
    def _align(ny, world):
        align = 64
        while align > 1 and ny % (align * world) != 0:
            align //= 2
        return align

    def step(self, dt):
        lib = _native.lib()
        d = (ctypes.c_double * 1)(float(dt))
        _native.check(lib.ldc_solver_set_dt(self._handle, d), 'set_dt')
        rep = _native.LdcNewtonReport()
        du = (ctypes.c_double * 1)(0.0)
        _native.check(lib.ldc_solver_step(self._handle, None, ctypes.byref(rep), du, _native.current_stream()),
                      'ldc_solver_step')
        self.newton_iterations += rep.iterations
        self.krylov_iterations += rep.linear_iterations
        return rep, du[0]

    def run(self, final_time=None, minimum_dt=1e-6, adaptative_dt=True, max_steps=None, on_step=None):
        """time_stepper.run_simulation for the distributed cavity (identical on every rank)"""
        t, dt = 0.0, self.dt
        while final_time is None or t < final_time:
            if final_time is not None and t + dt > final_time:
                dt = final_time - t
            rep, du = self.step(dt)
            if rep.reason <= 0:
                self.retries += 1
                if not adaptative_dt:
                    raise SolverDivergedException()
                dt /= 2.0
                if dt <= minimum_dt:
                    raise RuntimeError("Timestep has reached a too low value. Giving up.")
                continue
            t += dt
            self.steps += 1
            if on_step is not None:
                on_step(t, dt, rep, du)
            if final_time is None:
                if du / self.bc < 1e-6:
                    break
                if adaptative_dt:
                    dt *= 1.2
            elif abs(final_time - t) <= 1e-5 and adaptative_dt:
                dt *= 2.0
            if max_steps is not None and self.steps >= max_steps:
                break
        self.dt, self.t = dt, t
        return t
