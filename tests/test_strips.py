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
