"""Reynolds-number ensembles: many independent cavities of the same grid advanced in lock step
by ONE batched device solver (BASELINE.json config 4: 512 cavities of 64x64, Re 100..10000).

The reference has no such driver -- ``experiments/benchmarks.py:59-75`` loops over cases
serially, one ``run_simulation`` each.  Here every kernel carries a batch dimension, per-cavity
``bc`` and ``dt`` live in device arrays, and the time-stepper rules of ``time_stepper.py:9-56``
(halve dt on divergence, x1.2 growth, steady test, fixed-time clamp) are applied PER CAVITY on
the host from a few bytes of per-cavity status; converged / failed cavities are masked out on
the device.  Cavities are independent units, so several GPUs simply take contiguous slices of
the ensemble (``partition``) with no data-path collective; only the per-cavity summaries are
gathered.
"""
import ctypes
import logging

import numpy as np

logger = logging.getLogger(__name__)

RUNNING, STEADY, FINISHED, FAILED = 0, 1, 2, 3


def partition(n_items, world_size, rank):
    """Contiguous, balanced slice of ``range(n_items)`` owned by ``rank`` (first ranks get the
    remainder), e.g. 512 cavities over 8 GPUs -> 64 each."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


class EnsembleTimeStepper(object):
    """Host-side, per-cavity restatement of the reference time stepper (time_stepper.py:9-56).
    Pure numpy: it only sees per-cavity convergence reasons and steady-state norms, so it is
    testable without a GPU."""

    def __init__(self, dt0, bc, final_time=None, minimum_dt=1e-6, adaptative_dt=True):
        self.dt = np.array(dt0, dtype=np.float64).copy()
        self.bc = np.array(bc, dtype=np.float64).copy()
        self.t = np.zeros_like(self.dt)
        self.status = np.full(self.dt.shape, RUNNING, dtype=np.int32)
        self.steps = np.zeros(self.dt.shape, dtype=np.int64)
        self.retries = np.zeros(self.dt.shape, dtype=np.int64)
        self.final_time = final_time
        self.minimum_dt = minimum_dt
        self.adaptative_dt = adaptative_dt

    def active(self):
        return self.status == RUNNING

    def begin_step(self):
        """dt each active cavity will use (clamped onto final_time), time_stepper.py:16-18"""
        if self.final_time is not None:
            act = self.active()
            over = act & (self.t + self.dt > self.final_time)
            self.dt[over] = self.final_time - self.t[over]
        return self.dt

    def end_step(self, reasons, du_inf):
        """apply the stepper rules to every active cavity given the Newton reasons (>0 converged)
        and max|U-U_old| of the step just taken"""
        reasons = np.asarray(reasons)
        du_inf = np.asarray(du_inf, dtype=np.float64)
        for b in np.nonzero(self.active())[0]:
            if reasons[b] <= 0:                         # SolverDivergedException branch (:23-32)
                self.retries[b] += 1
                if self.adaptative_dt:
                    self.dt[b] /= 2.0
                if self.dt[b] <= self.minimum_dt or not self.adaptative_dt:
                    self.status[b] = FAILED
                continue
            self.t[b] += self.dt[b]
            self.steps[b] += 1
            if self.final_time is None:
                if du_inf[b] / self.bc[b] < 1e-6:      # steady test (:39-42)
                    self.status[b] = STEADY
                elif self.adaptative_dt:
                    self.dt[b] *= 1.2
            elif abs(self.final_time - self.t[b]) <= 1e-5:
                if self.adaptative_dt:
                    self.dt[b] *= 2.0
                self.status[b] = FINISHED
            elif self.t[b] >= self.final_time:
                self.status[b] = FINISHED


class CudaEnsembleSolver(object):
    """Batched device solver for ``graphs`` (same nx, ny, dx, dy, rho, mi; individual bc, dt and
    initial fields)."""

    def __init__(self, graphs, **solver_options):
        from lid_driven_cavity_problem_b200 import _native
        self._native = _native
        self.graphs = list(graphs)
        g0 = self.graphs[0]
        for g in self.graphs:
            same = (g.pressure_mesh.nx == g0.pressure_mesh.nx and g.pressure_mesh.ny == g0.pressure_mesh.ny and
                    g.dx == g0.dx and g.dy == g0.dy and g.rho == g0.rho and g.mi == g0.mi and
                    bool(getattr(g, 'use_cds', False)) == bool(getattr(g0, 'use_cds', False)))
            if not same:
                raise ValueError("all cavities of an ensemble must share grid, spacing, rho, mi and scheme")
        torch = _native.require_cuda()
        lib = _native.lib()
        self.batch = len(self.graphs)
        self.nx, self.ny = int(g0.pressure_mesh.nx), int(g0.pressure_mesh.ny)
        opt = _native.LdcSolverOptions()
        lib.ldc_solver_default_options(ctypes.byref(opt))
        for k, v in solver_options.items():
            if not hasattr(opt, k):
                raise TypeError("unknown solver option %r" % k)
            setattr(opt, k, v)
        grid = _native.grid_from_graph(g0, batch=self.batch)
        self._handle = ctypes.c_void_p()
        _native.check(lib.ldc_solver_create(ctypes.byref(grid), ctypes.byref(opt), ctypes.byref(self._handle)),
                      'ldc_solver_create')
        bc = (ctypes.c_double * self.batch)(*[float(g.bc) for g in self.graphs])
        _native.check(lib.ldc_solver_set_bc(self._handle, bc), 'set_bc')
        # upload the initial fields
        from lid_driven_cavity_problem_b200.nonlinear_solver._common import create_X_device
        U = torch.from_numpy(np.concatenate([np.asarray(g.ns_x_mesh.phi, dtype=np.float64) for g in self.graphs])).cuda()
        V = torch.from_numpy(np.concatenate([np.asarray(g.ns_y_mesh.phi, dtype=np.float64) for g in self.graphs])).cuda()
        P = torch.from_numpy(np.concatenate([np.asarray(g.pressure_mesh.phi, dtype=np.float64) for g in self.graphs])).cuda()
        X = create_X_device(U, V, P, g0, batch=self.batch)
        _native.check(lib.ldc_solver_set_state(self._handle, _native.ptr(X), _native.current_stream()), 'set_state')
        self.newton_iterations = np.zeros(self.batch, dtype=np.int64)
        self.krylov_iterations = np.zeros(self.batch, dtype=np.int64)

    def close(self):
        if self._handle is not None:
            self._native.lib().ldc_solver_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, dt, active):
        """one pseudo-time step of the active cavities; returns (reasons, du_inf) numpy arrays"""
        _native, lib = self._native, self._native.lib()
        B = self.batch
        dt_arr = (ctypes.c_double * B)(*[float(v) for v in dt])
        _native.check(lib.ldc_solver_set_dt(self._handle, dt_arr), 'set_dt')
        act = (ctypes.c_int32 * B)(*[int(bool(a)) for a in active])
        reports = (_native.LdcNewtonReport * B)()
        du = (ctypes.c_double * B)()
        _native.check(lib.ldc_solver_step(self._handle, act, reports, du, _native.current_stream()), 'ldc_solver_step')
        reasons = np.array([reports[b].reason for b in range(B)])
        for b in range(B):
            if active[b]:
                self.newton_iterations[b] += reports[b].iterations
                self.krylov_iterations[b] += reports[b].linear_iterations
        return reasons, np.array([du[b] for b in range(B)])

    def counters(self):
        out = (ctypes.c_int64 * 3)()
        self._native.check(self._native.lib().ldc_solver_counters(self._handle, out), 'counters')
        return dict(lockstep_krylov=int(out[0]), krylov_sum=int(out[1]), lockstep_newton=int(out[2]))

    def fields(self):
        """(U, V, P) as numpy arrays of shape (batch, mesh size)"""
        _native, lib = self._native, self._native.lib()
        import torch
        from lid_driven_cavity_problem_b200.nonlinear_solver._common import recover_X_device
        n = 3 * self.nx * self.ny * self.batch
        X = torch.empty(n, dtype=torch.float64, device='cuda')
        _native.check(lib.ldc_solver_get_state(self._handle, _native.ptr(X), _native.current_stream()), 'get_state')
        U, V, P = recover_X_device(X, self.graphs[0], batch=self.batch)
        return (U.cpu().numpy().reshape(self.batch, -1), V.cpu().numpy().reshape(self.batch, -1),
                P.cpu().numpy().reshape(self.batch, -1))

    def run(self, final_time=None, minimum_dt=1e-6, adaptative_dt=True, max_steps=None):
        ts = EnsembleTimeStepper([g.dt for g in self.graphs], [g.bc for g in self.graphs], final_time,
                                 minimum_dt, adaptative_dt)
        sweeps = 0
        while ts.active().any():
            dt = ts.begin_step()
            reasons, du = self.step(dt, ts.active())
            ts.end_step(reasons, du)
            sweeps += 1
            logger.info("ensemble sweep %d: %d running, %d steady, %d failed", sweeps,
                        int((ts.status == RUNNING).sum()), int((ts.status == STEADY).sum()),
                        int((ts.status == FAILED).sum()))
            if max_steps is not None and sweeps >= max_steps:
                break
        self.stepper = ts
        return ts


def interleaved(n_items, world_size, rank):
    """indices rank, rank+world, ...: when cavities are sorted by difficulty (Reynolds number) every
    rank gets the same mix, which balances the lock-step batches"""
    return list(range(rank, n_items, world_size))


def _solve_group(graphs, indices, rank, final_time, minimum_dt, adaptative_dt, max_steps, solver_options):
    """one batched solver for ``graphs`` (their global indices in ``indices``); returns
    (summaries, (U, V, P), counters)"""
    solver = CudaEnsembleSolver(graphs, **solver_options)
    try:
        ts = solver.run(final_time, minimum_dt, adaptative_dt, max_steps)
        fields = solver.fields()
        counters = dict(solver.counters() if hasattr(solver, 'counters') else {},
                        sweeps=int((ts.steps + ts.retries).max()))
        summaries = [dict(index=indices[b], bc=float(graphs[b].bc), status=int(ts.status[b]), t=float(ts.t[b]),
                          dt=float(ts.dt[b]), steps=int(ts.steps[b]), retries=int(ts.retries[b]),
                          newton=int(solver.newton_iterations[b]), krylov=int(solver.krylov_iterations[b]), rank=rank)
                     for b in range(len(graphs))]
    finally:
        solver.close()
    return summaries, fields, counters


def run_ensemble(graphs, final_time=None, minimum_dt=1e-6, adaptative_dt=True, max_steps=None,
                 distributed=False, interleave=False, groups=1, **solver_options):
    """Solve every cavity of ``graphs``.  With ``distributed=True`` (torch.distributed initialised,
    one process per GPU) each rank solves ``partition(len(graphs), world, rank)`` and the per-cavity
    summaries are gathered on every rank; there is no collective on the data path.

    ``groups`` > 1: the rank's cavities are sorted by lid speed (= Reynolds number = difficulty) and
    cut into that many sub-batches, each with its own batched solver, CUDA stream and host thread,
    all running CONCURRENTLY on the GPU.  A batch advances in lock step at the pace of its hardest
    cavity, and the hard cavities are one long latency-bound chain of small launches (one 64x64
    cavity at Re=10^4: 19 000 Krylov iterations of ~230 us, 4.5 s, with the GPU almost idle); grouping
    by difficulty lets the easy sub-batches finish early and the chains of the hard ones overlap
    instead of being summed.

    Returns a list of dicts (one per cavity, global order) and this rank's (U, V, P) in the order of
    its cavities."""
    rank, world = 0, 1
    if distributed:
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
    if interleave:
        mine_idx = interleaved(len(graphs), world, rank)
    else:
        sl = partition(len(graphs), world, rank)
        mine_idx = list(range(len(graphs))[sl])
    summaries, fields = [], None
    if len(mine_idx) > 0:
        groups = max(1, min(int(groups), len(mine_idx)))
        args = (rank, final_time, minimum_dt, adaptative_dt, max_steps, solver_options)
        if groups == 1:
            summaries, fields, counters = _solve_group([graphs[i] for i in mine_idx], mine_idx, *args)
            run_ensemble.last_counters = counters
        else:
            import threading
            import torch
            order = sorted(range(len(mine_idx)), key=lambda k: float(graphs[mine_idx[k]].bc))
            cuts = [len(order) * k // groups for k in range(groups + 1)]
            chunks = [order[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
            results, errors = [None] * len(chunks), [None] * len(chunks)
            on_gpu = torch.cuda.is_available()
            device = torch.cuda.current_device() if on_gpu else None

            def work(c):
                try:
                    idx = [mine_idx[k] for k in chunks[c]]
                    if on_gpu:
                        torch.cuda.set_device(device)
                        with torch.cuda.stream(torch.cuda.Stream()):
                            results[c] = _solve_group([graphs[i] for i in idx], idx, *args)
                            torch.cuda.current_stream().synchronize()
                    else:                       # host-logic tests with a fake solver
                        results[c] = _solve_group([graphs[i] for i in idx], idx, *args)
                except BaseException as e:      # noqa: B902  (re-raised on the calling thread)
                    errors[c] = e
            threads = [threading.Thread(target=work, args=(c,)) for c in range(len(chunks))]
            for t in reversed(threads):          # the hardest sub-batch first: it is the critical path
                t.start()
            for t in threads:
                t.join()
            for e in errors:
                if e is not None:
                    raise e
            # back to the order of mine_idx
            pos = {}
            for c, chunk in enumerate(chunks):
                for b, k in enumerate(chunk):
                    pos[k] = (c, b)
            summaries = [results[pos[k][0]][0][pos[k][1]] for k in range(len(mine_idx))]
            if all(r[1] is not None for r in results):
                fields = tuple(np.stack([results[pos[k][0]][1][f][pos[k][1]] for k in range(len(mine_idx))])
                               for f in range(3))
            run_ensemble.last_counters = dict(groups=[r[2] for r in results])
    if distributed:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, summaries)
        summaries = sorted((s for part in gathered for s in part), key=lambda d: d['index'])
    return summaries, fields
