"""Test infrastructure: an in-process implementation of the ``ldc_comm_ops`` function table
(include/ldc_b200.h) that lets several row strips of ONE cavity run in ONE process on ONE GPU --
one Python thread per strip, a ``threading.Barrier`` as the rendezvous, device-to-device copies as
the transport.  It exists so that the distributed code path of the solver (halo rows of fp64 and
fp32 level vectors, all-reduced norms, the all-gather into the replicated coarse levels) is
exercised by ``pytest -m gpu`` on a single-GPU box; the NCCL transport
(lid_driven_cavity_problem_b200/csrc/comm.cu) implements the same three calls.

Every call synchronises the device and meets the other strips at a barrier, so solvers using it
must be created with ``use_cuda_graphs=0`` (a host rendezvous cannot be captured)."""
import ctypes
import threading

import torch

from lid_driven_cavity_problem_b200 import peer_halo

HALO_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                           ctypes.c_void_p)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                ctypes.c_void_p)
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)


class LdcCommOps(ctypes.Structure):
    _fields_ = [('ctx', ctypes.c_void_p), ('rank', ctypes.c_int32), ('nranks', ctypes.c_int32),
                ('halo_exchange', HALO_FN), ('allreduce', ALLREDUCE_FN), ('allgather', ALLGATHER_FN)]


def _span(ptr, doubles):
    return torch.as_tensor(peer_halo._DeviceSpan(int(ptr), int(doubles)), device='cuda')


class ThreadTransport(object):
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.calls = {'halo': 0, 'allreduce': 0, 'allgather': 0}
        self._keep = []

    def _guard(self, fn):
        def wrapped(*a):
            try:
                fn(*a)
                return 0
            except Exception:                     # never leave the other strips waiting
                import traceback
                traceback.print_exc()
                self.barrier.abort()
                return 1
        return wrapped

    def comm(self, rank):
        """object with the ``ops()`` method CudaStripSolver expects"""
        world, slots, barrier = self.world, self.slots, self.barrier

        def halo(ctx, first, row, rows, stream):
            torch.cuda.synchronize()
            slots[rank] = (first, row, rows)
            barrier.wait()
            if rank > 0:
                pf, prow, prows = slots[rank - 1]
                assert prow == row
                _span(first - 8 * row, row).copy_(_span(pf + 8 * prow * (prows - 1), prow))
            if rank < world - 1:
                nf, nrow, _ = slots[rank + 1]
                assert nrow == row
                _span(first + 8 * row * rows, row).copy_(_span(nf, nrow))
            torch.cuda.synchronize()
            barrier.wait()
            if rank == 0:
                self.calls['halo'] += 1

        def allreduce(ctx, buf, count, op, stream):
            torch.cuda.synchronize()
            slots[rank] = (buf, count)
            barrier.wait()
            parts = torch.stack([_span(b, c).clone() for b, c in slots])
            assert all(c == count for _, c in slots)
            torch.cuda.synchronize()
            barrier.wait()                        # everybody has read everybody's input
            _span(buf, count).copy_(parts.max(dim=0).values if op == 1 else parts.sum(dim=0))
            torch.cuda.synchronize()
            barrier.wait()
            if rank == 0:
                self.calls['allreduce'] += 1

        def allgather(ctx, buf, count, stream):
            torch.cuda.synchronize()
            slots[rank] = (buf, count)
            barrier.wait()
            for r, (b, c) in enumerate(slots):
                assert c == count
                if r != rank:
                    _span(buf + 8 * count * r, count).copy_(_span(b + 8 * count * r, count))
            torch.cuda.synchronize()
            barrier.wait()
            if rank == 0:
                self.calls['allgather'] += 1

        table = LdcCommOps()
        table.ctx, table.rank, table.nranks = None, rank, world
        table.halo_exchange = HALO_FN(self._guard(halo))
        table.allreduce = ALLREDUCE_FN(self._guard(allreduce))
        table.allgather = ALLGATHER_FN(self._guard(allgather))
        self._keep.append(table)

        class _Comm(object):
            def ops(self):
                return ctypes.cast(ctypes.pointer(table), ctypes.c_void_p)
        return _Comm()


def run_strips(world, make_and_run):
    """``make_and_run(rank, comm)`` in one thread per strip; returns the list of results, re-raises
    the first failure"""
    transport = ThreadTransport(world)
    out, err = [None] * world, [None] * world

    def work(rank):
        try:
            torch.cuda.set_device(0)
            out[rank] = make_and_run(rank, transport.comm(rank))
        except BaseException as e:                # noqa: B902
            err[rank] = e
            transport.barrier.abort()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for e in err:
        if e is not None:
            raise e
    assert all(not t.is_alive() for t in threads), "a strip thread is stuck"
    return out, transport
