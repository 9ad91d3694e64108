"""Row-strip residual without a collective (``ldc_residual_peer_push`` + ``ldc_residual_peer``,
include/ldc_b200.h): every strip pushes its boundary rows into the neighbours' ghost rows over
NVLink (a small kernel on a side stream) and publishes the evaluation number in their flag
words; the residual kernel's boundary tiles wait for their own words and read the ghost rows
from local memory.  One evaluation is one residual launch plus one overlapped push.

The reference has no multi-rank path (sequential PETSc objects only,
nonlinear_solver/petsc_solver_wrapper.py:75,77,117,125); this is the SURVEY 8(e) strip
decomposition of ``residual_function(X, graph)`` (c++/residual_function.cpp:10-238).

Region layout (identical on every strip, so a neighbour's addresses follow from its base):

    [0, 16)     two 64-bit flag words: [0] written by strip r-1, [1] by strip r+1
    [64, 68)    status word (1 = a wait timed out)
    [256, ...)  ``n_sets`` X buffers, each ``3*nx*(rows_max+2)`` doubles:
                [ghost row below | owned rows | ghost row above | ...]; a strip owns
                ``rows`` <= ``rows_max`` rows
"""
import ctypes

from lid_driven_cavity_problem_b200 import _native

HEADER_BYTES = 256
STATUS_OFFSET = 64


class _DeviceSpan(object):
    """minimal __cuda_array_interface__ carrier so torch can view memory the C library owns"""

    def __init__(self, address, n_doubles):
        self.__cuda_array_interface__ = {'shape': (int(n_doubles),), 'typestr': '<f8',
                                         'data': (int(address), False), 'version': 2}


def region_bytes(nx, rows_max, n_sets):
    return HEADER_BYTES + 8 * 3 * nx * (rows_max + 2) * n_sets


class PeerStripResidual(object):
    """One strip.  ``local``/``prev``/``next``: base addresses of this strip's region and of the
    neighbours' regions as seen from this process (``Comm.peer_alloc`` across ranks, or plain
    device buffers of one process when several strips share a GPU); ``prev_rows``: rows owned by
    strip r-1."""

    def __init__(self, grid, local, prev, next, prev_rows, rows_max, n_sets=1):
        self.lib = _native.lib()
        self.grid = grid
        self.nx, self.rows = int(grid.nx), int(grid.row_count)
        self.row_bytes = 8 * 3 * self.nx
        self.set_bytes = self.row_bytes * (rows_max + 2)
        self.local, self.prev, self.next = int(local), int(prev or 0), int(next or 0)
        self.prev_rows, self.n_sets = int(prev_rows), int(n_sets)
        self.epoch = 0
        self._ctx = None
        has_prev = grid.row_begin > 0
        has_next = grid.row_begin + grid.row_count < grid.ny
        if has_prev and not self.prev or has_next and not self.next:
            raise ValueError("neighbour region missing")
        self._has_prev, self._has_next = has_prev, has_next

    def x_address(self, k=0):
        """device address of the first owned row of X buffer ``k``"""
        return self.local + HEADER_BYTES + k * self.set_bytes + self.row_bytes

    def x_tensor(self, k=0):
        """torch view (3*nx*rows doubles) of the owned rows of X buffer ``k``"""
        import torch
        return torch.as_tensor(_DeviceSpan(self.x_address(k), 3 * self.nx * self.rows), device='cuda')

    def status(self):
        """True when any wait of this strip timed out (synchronises the device)"""
        import torch
        torch.cuda.synchronize()
        t = torch.as_tensor(_DeviceSpan(self.local + STATUS_OFFSET, 1), device='cuda')
        return bool(t.view(torch.int64)[0].item() & 0xffffffff)

    def _peer_struct(self, k):
        ph = _native.LdcPeerHalo()
        off = HEADER_BYTES + k * self.set_bytes
        if self._has_prev:
            ph.prev_ghost = self.prev + off + self.row_bytes * (self.prev_rows + 1)   # above its last owned row
            ph.prev_flag = self.prev + 8                                              # its my_flags[1]
        if self._has_next:
            ph.next_ghost = self.next + off                                           # below its first owned row
            ph.next_flag = self.next                                                  # its my_flags[0]
        ph.my_flags = self.local
        ph.status_dev = self.local + STATUS_OFFSET
        return ph

    def bind_overlapped(self, k, U_old, V_old, F, norm2=None, work=None, lag=4):
        """``launch(stream)``: ONE library call per evaluation (ldc_residual_peer_step) that puts the
        push and the strip's edge tiles on the library's two high-priority streams and the rows in
        between on ``stream``; ``join(stream)`` before F is consumed.  For loops over rotating X
        buffers that are final before the loop starts; needs ``n_sets >= 2*lag + 3``
        (see include/ldc_b200.h)."""
        if self.n_sets < 2 * lag + 3:
            raise ValueError("overlapped evaluation needs at least 2*lag+3 rotating X buffers")
        if self._ctx is None:
            self._ctx = ctypes.c_void_p()
            _native.check(self.lib.ldc_peer_ctx_create(lag, ctypes.byref(self._ctx)), 'ldc_peer_ctx_create')
        ph = self._peer_struct(k)
        ph_ref = ctypes.byref(ph)
        fn, grid, ctx = self.lib.ldc_residual_peer_step, self.grid, self._ctx
        args = (ctypes.c_void_p(self.x_address(k)), _native.ptr(U_old), _native.ptr(V_old), _native.ptr(F),
                _native.ptr(norm2), _native.ptr(work))

        def launch(stream):
            self.epoch += 1
            ph.epoch = self.epoch
            rc = fn(ctx, grid, ph_ref, args[0], args[1], args[2], args[3], args[4], args[5], stream)
            if rc != 0:
                _native.check(rc, 'ldc_residual_peer_step')
        launch.keep = (U_old, V_old, F, norm2, work)
        return launch

    def join(self, stream):
        """make ``stream`` wait for the side-stream part (edge tiles) of the last overlapped
        evaluation: F of that evaluation is complete on ``stream`` afterwards"""
        if self._ctx is not None:
            _native.check(self.lib.ldc_peer_ctx_join(self._ctx, stream), 'ldc_peer_ctx_join')

    def close(self):
        if self._ctx is not None:
            self.lib.ldc_peer_ctx_destroy(self._ctx)
            self._ctx = None

    def residual(self, k, U_old, V_old, F, stream=None, norm2=None, work=None, push_stream=None):
        """F <- residual of the strip's owned rows of X buffer ``k`` (push + kernel); every strip must
        make the same sequence of calls (the evaluation number `epoch` is what the flag words carry)."""
        self.bind(k, U_old, V_old, F, norm2, work)(stream if stream is not None else _native.current_stream(),
                                                   push_stream)

    def bind(self, k, U_old, V_old, F, norm2=None, work=None):
        """``launch(stream, push_stream=None, push=True, kernel=True)`` for a fixed set of buffers
        with every argument converted once (the kernel runs for ~14 us, a hot loop should not
        rebuild ctypes structures per call).  ``push_stream``: where the halo push goes (``None`` =
        the launch stream, i.e. serialised in front of the kernel).  X of buffer ``k`` must be final
        on that stream; the caller also keeps pushes from running more than ``n_sets - 2``
        evaluations ahead of the neighbours (see include/ldc_b200.h)."""
        ph = self._peer_struct(k)
        ph_ref = ctypes.byref(ph)
        fn, push_fn, grid = self.lib.ldc_residual_peer, self.lib.ldc_residual_peer_push, self.grid
        x_ptr = ctypes.c_void_p(self.x_address(k))
        args = (x_ptr, _native.ptr(U_old), _native.ptr(V_old), _native.ptr(F), _native.ptr(norm2), _native.ptr(work))
        keep = (U_old, V_old, F, norm2, work)
        has_neighbour = self._has_prev or self._has_next

        def launch(stream, push_stream=None, push=True, kernel=True):
            if push:
                self.epoch += 1
            ph.epoch = self.epoch
            if push and has_neighbour:
                rc = push_fn(grid, ph_ref, x_ptr, push_stream if push_stream is not None else stream)
                if rc != 0:
                    _native.check(rc, 'ldc_residual_peer_push')
            if kernel:
                rc = fn(grid, ph_ref, args[0], args[1], args[2], args[3], args[4], args[5], stream)
                if rc != 0:
                    _native.check(rc, 'ldc_residual_peer')
        launch.keep = keep
        return launch
