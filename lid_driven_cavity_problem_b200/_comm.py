"""ctypes binding of include/ldc_b200_comm.h (libldc_b200_comm.so, NCCL transport of the row-strip
decomposition) and the host-side partitioning helpers.  torch.distributed is only the bootstrap
channel for the NCCL unique id; the data path calls NCCL from C."""
import ctypes
import os

from lid_driven_cavity_problem_b200 import _native

_HERE = os.path.dirname(os.path.abspath(__file__))
COMM_LIB_PATH = os.path.join(_HERE, 'csrc', 'libldc_b200_comm.so')
ID_BYTES = 128

_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
COMM_SYMBOLS = {
    'ldc_comm_unique_id': (ctypes.c_int, [_vp]),
    'ldc_comm_create': (ctypes.c_int, [_vp, _i32, _i32, ctypes.POINTER(_vp)]),
    'ldc_comm_destroy': (None, [_vp]),
    'ldc_comm_last_error': (ctypes.c_char_p, []),
    'ldc_comm_halo_exchange': (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp]),
    'ldc_comm_allreduce': (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    'ldc_comm_allgather': (ctypes.c_int, [_vp, _vp, _i64, _vp]),
    'ldc_comm_peer_alloc': (ctypes.c_int, [_vp, _i64, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    'ldc_comm_peer_free': (ctypes.c_int, [_vp, _vp]),
    'ldc_comm_get_ops': (_vp, [_vp]),
}
_lib = None


def load_comm_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(COMM_LIB_PATH):
        raise RuntimeError("CUDA comm module is not compiled. (%s missing)" % COMM_LIB_PATH)
    import torch  # noqa: F401  (loads the libnccl.so.2 the comm library is linked against)
    lib = ctypes.CDLL(COMM_LIB_PATH)
    for name, (restype, argtypes) in COMM_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load_comm_library().ldc_comm_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else '?'))


def strip_partition(ny, world_size, rank, align=1):
    """(row_begin, row_count) of the row strip owned by ``rank``: contiguous, balanced, and -- when
    possible -- every boundary a multiple of ``align`` rows (a power of two keeps the strips
    coarsenable for the multigrid hierarchy)."""
    if ny % (align * world_size) != 0:
        align = 1
    units = ny // align
    base, rem = divmod(units, world_size)
    start = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return start * align, count * align


class Comm(object):
    """One NCCL communicator over the ranks of the default torch.distributed group."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        lib = load_comm_library()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        ident = [None]
        if self.rank == 0:
            buf = ctypes.create_string_buffer(ID_BYTES)
            check(lib.ldc_comm_unique_id(buf), 'ldc_comm_unique_id')
            ident[0] = bytes(buf.raw)
        dist.broadcast_object_list(ident, src=0)
        self._handle = ctypes.c_void_p()
        check(lib.ldc_comm_create(ident[0], self.world, self.rank, ctypes.byref(self._handle)), 'ldc_comm_create')
        self._lib = lib
        torch.cuda.synchronize()

    def ops(self):
        """pointer to the ldc_comm_ops table (for ldc_solver_set_comm)"""
        return ctypes.c_void_p(self._lib.ldc_comm_get_ops(self._handle))

    def halo_exchange(self, first_owned_row_ptr, row_doubles, row_count, stream=None):
        check(self._lib.ldc_comm_halo_exchange(self._handle, first_owned_row_ptr, row_doubles, row_count,
                                               stream if stream is not None else _native.current_stream()),
              'ldc_comm_halo_exchange')

    def allreduce(self, tensor, op='sum', stream=None):
        check(self._lib.ldc_comm_allreduce(self._handle, tensor.data_ptr(), tensor.numel(),
                                           1 if op == 'max' else 0,
                                           stream if stream is not None else _native.current_stream()),
              'ldc_comm_allreduce')

    def peer_alloc(self, nbytes):
        """collective: (local, prev, next) device addresses of a zero-filled region of ``nbytes`` on
        this rank and of the neighbour ranks' regions mapped into this process (0 where none)"""
        loc, prv, nxt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        check(self._lib.ldc_comm_peer_alloc(self._handle, int(nbytes), ctypes.byref(loc), ctypes.byref(prv),
                                            ctypes.byref(nxt)), 'ldc_comm_peer_alloc')
        return loc.value or 0, prv.value or 0, nxt.value or 0

    def peer_free(self, local):
        check(self._lib.ldc_comm_peer_free(self._handle, ctypes.c_void_p(local)), 'ldc_comm_peer_free')

    def close(self):
        if self._handle is not None:
            self._lib.ldc_comm_destroy(self._handle)
            self._handle = None
