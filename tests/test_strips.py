"""Row-strip decomposition: host-side partitioning (CPU), the NCCL transport library's ABI (CPU,
no communication), a 2-rank gloo run of the strip bookkeeping, and the strip code path of the
device solver on one GPU."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO


def test_strip_partition_covers_the_grid_and_stays_coarsenable():
    from lid_driven_cavity_problem_b200._comm import strip_partition
    for ny, world, align in [(4096, 8, 64), (4096, 4, 64), (1024, 2, 64), (100, 3, 64), (256, 1, 64), (70, 4, 1)]:
        parts = [strip_partition(ny, world, r, align) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == ny
        for (b0, c0), (b1, _) in zip(parts[:-1], parts[1:]):
            assert b0 + c0 == b1
        if ny % (align * world) == 0:
            assert all(b % align == 0 and c % align == 0 for b, c in parts)
        assert max(c for _, c in parts) - min(c for _, c in parts) <= (align if ny % (align * world) == 0 else 1)


def _plan_levels(nx, ny, row_begin, row_count):
    import ctypes
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.lib()
    g = _native.LdcGrid()
    g.nx, g.ny, g.row_begin, g.row_count, g.batch = nx, ny, row_begin, row_count, 1
    g.dt, g.dx, g.dy, g.rho, g.mi, g.bc = 1e-2, 1.0 / nx, 1.0 / ny, 1.0, 1.0, 1.0
    opt = _native.LdcSolverOptions()
    lib.ldc_solver_default_options(ctypes.byref(opt))
    lev, rep = ctypes.c_int32(), ctypes.c_int32()
    _native.check(lib.ldc_solver_plan_levels(ctypes.byref(g), ctypes.byref(opt), ctypes.byref(lev),
                                             ctypes.byref(rep)), 'ldc_solver_plan_levels')
    return lev.value, rep.value


def test_every_rank_of_an_equal_partition_plans_the_same_hierarchy():
    """The multigrid shape is computed per rank from its own row_begin / row_count
    (ldc_solver_plan_levels, host arithmetic).  Ranks that disagree would issue different
    sequences of collectives and hang, so (a) equal strips -- what CudaStripSolver hands out --
    must agree on every rank, (b) an unequal split is refused on every rank before any device or
    NCCL work (ValueError here; ldc_solver_set_comm repeats the check collectively)."""
    from lid_driven_cavity_problem_b200._comm import strip_partition
    from lid_driven_cavity_problem_b200.strip_solver import CudaStripSolver
    for nx, ny, world in [(4096, 4096, 8), (4096, 4096, 4), (4096, 4096, 2), (1024, 1024, 8), (1000, 1000, 4),
                          (1000, 1000, 5), (96, 72, 3), (64, 66, 2), (512, 384, 6), (256, 256, 1)]:
        assert ny % world == 0
        align = CudaStripSolver._align(ny, world)
        plans = {_plan_levels(nx, ny, *strip_partition(ny, world, r, align)) for r in range(world)}
        assert len(plans) == 1, (nx, ny, world, plans)
    assert _plan_levels(4096, 4096, 512, 512)[0] >= 6       # a real hierarchy, replicated below some level
    assert _plan_levels(4096, 4096, 512, 512)[1] > 0
    # the advisor's example: 1000 rows over 3 ranks = 334/333/333 -> the ranks would disagree ...
    uneven = {_plan_levels(1000, 1000, *strip_partition(1000, 3, r, 1)) for r in range(3)}
    assert len(uneven) > 1
    # ... so the host side refuses it identically on every rank, before touching a device
    for rank in range(3):
        with pytest.raises(ValueError, match="equal"):
            CudaStripSolver(1000, 1000, 1e-2, 1.0, 1.0, 1000.0, rank=rank, world=3)


def test_comm_library_exports_every_declared_symbol():
    from lid_driven_cavity_problem_b200 import _comm
    text = open(os.path.join(REPO, 'include', 'ldc_b200_comm.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    declared = sorted(set(re.findall(r'\b(ldc_comm_[A-Za-z0-9_]+)\s*\(', text)))
    lib = _comm.load_comm_library()
    assert len(declared) >= 7
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_comm.COMM_SYMBOLS) == set(declared)



def test_peer_halo_region_layout_and_handshake_addresses(monkeypatch):
    """host logic of peer_halo.PeerStripResidual (no GPU): where a strip looks for its neighbours'
    boundary rows and flag words inside their regions, and the epoch bookkeeping"""
    import ctypes
    from lid_driven_cavity_problem_b200 import _native, peer_halo
    nx, rows_max, n_sets = 12, 10, 3
    row_b = 8 * 3 * nx
    set_b = row_b * (rows_max + 2)
    assert peer_halo.region_bytes(nx, rows_max, n_sets) == peer_halo.HEADER_BYTES + n_sets * set_b
    g = _native.LdcGrid()
    g.nx, g.ny, g.row_begin, g.row_count, g.batch = nx, 27, 7, 10, 1
    calls = []

    class FakeLib(object):
        @staticmethod
        def ldc_residual_peer(grid, ph_ref, X, U, V, F, n2, work, stream):
            ph = ph_ref._obj
            calls.append(('kernel', ph.prev_ghost, ph.next_ghost, ph.my_flags, ph.prev_flag, ph.next_flag, ph.epoch,
                          ph.status_dev, X.value, stream.value))
            return 0

        @staticmethod
        def ldc_residual_peer_push(grid, ph_ref, X, stream):
            ph = ph_ref._obj
            calls.append(('push', ph.prev_ghost, ph.next_ghost, ph.my_flags, ph.prev_flag, ph.next_flag, ph.epoch,
                          ph.status_dev, X.value, stream.value))
            return 0
    local, prev, nxt = 0x100000, 0x200000, 0x300000
    monkeypatch.setattr(_native, 'lib', lambda: FakeLib)
    strip = peer_halo.PeerStripResidual(g, local, prev, nxt, prev_rows=7, rows_max=rows_max, n_sets=n_sets)
    launch = strip.bind(2, None, None, None)
    launch(ctypes.c_void_p(5))
    launch(ctypes.c_void_p(5), ctypes.c_void_p(9))      # push on a side stream
    off = peer_halo.HEADER_BYTES + 2 * set_b
    assert [c[0] for c in calls] == ['push', 'kernel', 'push', 'kernel']
    assert [c[9] for c in calls] == [5, 5, 9, 5]
    for kind, p_ghost, n_ghost, flags, p_flag, n_flag, epoch, status, x, _ in calls:
        assert p_ghost == prev + off + row_b * 8         # ghost slot + 7 owned rows -> the ghost row ABOVE them
        assert n_ghost == nxt + off                      # the ghost row BELOW its first owned row
        assert (flags, p_flag, n_flag) == (local, prev + 8, nxt)
        assert status == local + peer_halo.STATUS_OFFSET and x == local + off + row_b
    assert [c[6] for c in calls] == [1, 1, 2, 2]         # one epoch per evaluation, same for push and kernel
    # push and kernel can be issued separately (all pushes of an evaluation first)
    launch(ctypes.c_void_p(5), kernel=False)
    launch(ctypes.c_void_p(5), push=False)
    assert [c[0] for c in calls[-2:]] == ['push', 'kernel'] and [c[6] for c in calls[-2:]] == [3, 3]
    # first / last strips have one neighbour only; a missing region is refused
    g.row_begin, g.row_count = 0, 7
    first = peer_halo.PeerStripResidual(g, local, 0, nxt, 0, rows_max, n_sets)
    first.residual(0, None, None, None, ctypes.c_void_p(0))
    assert calls[-1][1] is None and calls[-1][4] is None and calls[-1][2] == nxt + peer_halo.HEADER_BYTES
    with pytest.raises(ValueError):
        peer_halo.PeerStripResidual(g, local, 0, 0, 0, rows_max, n_sets)


def test_overlapped_strip_step_host_contract(monkeypatch):
    """bind_overlapped: one ldc_residual_peer_step per evaluation with a strictly increasing epoch, a
    context created once, join() forwarded, and the rotation rule of the amortised flow control
    (at least 2*lag+3 X buffers, include/ldc_b200.h) enforced before anything is launched."""
    import ctypes
    from lid_driven_cavity_problem_b200 import _native, peer_halo
    nx, rows_max = 12, 10
    g = _native.LdcGrid()
    g.nx, g.ny, g.row_begin, g.row_count, g.batch = nx, 27, 7, 10, 1
    log = []

    class FakeLib(object):
        @staticmethod
        def ldc_peer_ctx_create(lag, out_ref):
            out_ref._obj.value = 0xC0FFEE
            log.append(('create', lag))
            return 0

        @staticmethod
        def ldc_residual_peer_step(ctx, grid, ph_ref, X, U, V, F, n2, work, stream):
            log.append(('step', ctx.value, ph_ref._obj.epoch, X.value, stream.value))
            return 0

        @staticmethod
        def ldc_peer_ctx_join(ctx, stream):
            log.append(('join', ctx.value, stream.value))
            return 0

        @staticmethod
        def ldc_peer_ctx_destroy(ctx):
            log.append(('destroy', ctx.value))
    monkeypatch.setattr(_native, 'lib', lambda: FakeLib)
    local, prev, nxt = 0x100000, 0x200000, 0x300000
    few = peer_halo.PeerStripResidual(g, local, prev, nxt, prev_rows=7, rows_max=rows_max, n_sets=8)
    with pytest.raises(ValueError, match='2\\*lag\\+3'):
        few.bind_overlapped(0, None, None, None, lag=3)
    assert log == []
    strip = peer_halo.PeerStripResidual(g, local, prev, nxt, prev_rows=7, rows_max=rows_max, n_sets=9)
    launches = [strip.bind_overlapped(k, None, None, None, lag=3) for k in range(9)]
    for k in (0, 1, 2, 0):
        launches[k](ctypes.c_void_p(5))
    strip.join(ctypes.c_void_p(5))
    strip.close()
    strip.close()
    assert log[0] == ('create', 3) and [e[0] for e in log].count('create') == 1
    steps = [e for e in log if e[0] == 'step']
    assert [e[2] for e in steps] == [1, 2, 3, 4] and all(e[1] == 0xC0FFEE and e[4] == 5 for e in steps)
    assert [e[3] for e in steps] == [strip.x_address(k) for k in (0, 1, 2, 0)]
    assert log[-2:] == [('join', 0xC0FFEE, 5), ('destroy', 0xC0FFEE)]


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["LDC_REPO"])
import numpy as np, torch
import torch.distributed as dist
from lid_driven_cavity_problem_b200._comm import strip_partition
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# the halo-exchange contract of include/ldc_b200_comm.h on CPU tensors: [ghost | rows | ghost]
nx, ny = 6, 8
rb, rc = strip_partition(ny, world, rank, 2)
row = 3 * nx
full = torch.arange(ny * row, dtype=torch.float64)
buf = torch.full(((rc + 2) * row,), -1.0, dtype=torch.float64)
buf[row:(rc + 1) * row] = full[rb * row:(rb + rc) * row]
ops = []
if rank > 0:
    ops += [dist.P2POp(dist.isend, buf[row:2 * row], rank - 1), dist.P2POp(dist.irecv, buf[0:row], rank - 1)]
if rank < world - 1:
    ops += [dist.P2POp(dist.isend, buf[rc * row:(rc + 1) * row], rank + 1),
            dist.P2POp(dist.irecv, buf[(rc + 1) * row:], rank + 1)]
for w in dist.batch_isend_irecv(ops):
    w.wait()
lo, hi = max(rb - 1, 0), min(rb + rc + 1, ny)
expect = full[lo * row:hi * row]
got = buf[(0 if rb > 0 else row):((rc + 2) * row if rb + rc < ny else (rc + 1) * row)]
assert torch.equal(got, expect), (rank, got, expect)
# all ranks take identical decisions from all-reduced scalars
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t)
assert t.item() == world * (world + 1) / 2
sys.stdout.write("rank%dok\n" % rank); sys.stdout.flush()
dist.barrier(); dist.destroy_process_group()
'''


def test_halo_layout_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                       env=dict(os.environ, LDC_REPO=REPO), capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'rank0ok' in r.stdout and 'rank1ok' in r.stdout


@pytest.mark.gpu
def test_strip_solver_single_rank_equals_cavity_solver():
    """world = 1: the strip code path (padded vectors, halo hooks, replicated-level logic idle) must
    give exactly the plain solver's answer.  (world >= 2 is checked on real GPUs by
    `tools/run_strips.py --check`, results in profiles/r1_multigpu.md.)"""
    from lid_driven_cavity_problem_b200.nonlinear_solver import cuda_solver_wrapper
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    from lid_driven_cavity_problem_b200.strip_solver import CudaStripSolver
    from lid_driven_cavity_problem_b200.time_stepper import run_simulation
    n, bc = 64, 400.0
    s = CudaStripSolver(n, n, 1e-2, 1.0, 1.0, bc)
    s.set_uniform_state()
    s.run(max_steps=4)
    X = s.gather_state().cpu().numpy()
    s.close()
    w = cuda_solver_wrapper.CudaSolverWrapper(None)
    ref = run_simulation(Graph(1.0, 1.0, n, n, 1e-2, 1.0, 1.0, bc), None, w.solve, max_steps=4)
    U = X[1::3].reshape(n, n)[:, :n - 1].ravel()
    assert np.array_equal(U, ref.ns_x_mesh.phi)
    assert np.array_equal(X[2::3][:n * (n - 1)], ref.ns_y_mesh.phi)
    assert s.krylov_iterations == w.total_linear_iterations


@pytest.mark.gpu
@pytest.mark.parametrize("world,fp32", [(2, 0), (2, 1), (4, 1), (8, 1)])
def test_distributed_strip_solve_equals_single_gpu_solve(world, fp32):
    """The DISTRIBUTED solver path on one GPU: `world` strips of one 128x128 cavity, one thread per
    strip, joined by the in-process transport of tests/thread_transport.py (same three calls as the
    NCCL library).  Exercises the halo rows of fp64 and -- with the mixed-precision cycle -- fp32
    level vectors, the all-reduced norms / Gram-Schmidt coefficients and the all-gather into the
    replicated coarse levels (8 strips: the replicated 64x64 level is larger than a strip's own 16
    fine rows).  With tight tolerances every Newton solve converges to ~1e-9, so the
    distributed and the single-GPU run must land on the same fields (1e-7 relative), whatever the
    summation order inside the preconditioner."""
    import torch
    from thread_transport import run_strips
    from lid_driven_cavity_problem_b200.strip_solver import CudaStripSolver
    n, bc, steps = 128, 400.0, 2
    tight = dict(snes_rtol=1e-13, snes_atol=1e-9, snes_stol=1e-15, ksp_rtol=1e-8, snes_max_it=30, gmres_refine=1,
                 pc_fp32=fp32)
    one = CudaStripSolver(n, n, 1e-2, 1.0, 1.0, bc, **tight)
    one.set_uniform_state()
    one.run(max_steps=steps)
    X1 = one.get_state().cpu().numpy()
    its1 = one.krylov_iterations
    one.close()

    def strip(rank, comm):
        s = CudaStripSolver(n, n, 1e-2, 1.0, 1.0, bc, comm=comm, rank=rank, world=world, use_cuda_graphs=0, **tight)
        try:
            s.set_uniform_state()
            s.run(max_steps=steps)
            torch.cuda.synchronize()
            return s.row_begin, s.row_count, s.get_state().cpu().numpy(), s.krylov_iterations, s.steps
        finally:
            s.close()
    parts, transport = run_strips(world, strip)
    assert [p[0] for p in parts] == [r * (n // world) for r in range(world)]
    X = np.concatenate([p[2] for p in parts])
    assert all(p[4] == steps for p in parts)
    assert len({p[3] for p in parts}) == 1                       # every rank took the same decisions
    assert abs(parts[0][3] - its1) <= max(3, its1 // 10), (parts[0][3], its1)
    assert transport.calls['halo'] > 0 and transport.calls['allreduce'] > 0 and transport.calls['allgather'] > 0
    scale = np.abs(X1[1::3]).max()
    assert np.abs(X[1::3] - X1[1::3]).max() < 1e-7 * scale
    assert np.abs(X[2::3] - X1[2::3]).max() < 1e-7 * scale
    P, P1 = X[0::3], X1[0::3]
    assert np.abs((P - P.mean()) - (P1 - P1.mean())).max() < 1e-7 * np.abs(P1 - P1.mean()).max()


@pytest.mark.gpu
def test_unequal_strip_hierarchies_are_refused_collectively():
    """ldc_solver_set_comm all-reduces the strip shapes: two strips of 96 and 32 rows of a 128-row
    cavity plan different hierarchies, and BOTH ranks get an error instead of a hang."""
    import ctypes
    from thread_transport import run_strips
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.lib()
    n = 128
    spans = [(0, 96), (96, 32)]

    def strip(rank, comm):
        g = _native.LdcGrid()
        g.nx, g.ny, g.row_begin, g.row_count, g.batch = n, n, spans[rank][0], spans[rank][1], 1
        g.dt, g.dx, g.dy, g.rho, g.mi, g.bc = 1e-2, 1.0 / n, 1.0 / n, 1.0, 1.0, 100.0
        h = ctypes.c_void_p()
        _native.check(lib.ldc_solver_create(ctypes.byref(g), None, ctypes.byref(h)))
        try:
            rc = lib.ldc_solver_set_comm(h, comm.ops())
            return rc, lib.ldc_last_error()
        finally:
            lib.ldc_solver_destroy(h)
    out, _ = run_strips(2, strip)
    assert all(rc != 0 for rc, _ in out)
    assert all(b'differ across ranks' in msg for _, msg in out)
