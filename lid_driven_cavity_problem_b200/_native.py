"""ctypes binding of the C-ABI in include/ldc_b200.h (libldc_b200.so, sm_100a).

PyTorch is only the owner of device memory and streams: every call passes
``tensor.data_ptr()`` and ``torch.cuda.current_stream().cuda_stream``.  There is no CPU
fallback: importing works without a GPU (so that the C-ABI can be inspected), but every
compute entry point raises if the library or a CUDA device is missing.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libldc_b200.so')

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class LdcGrid(ctypes.Structure):
    _fields_ = [('nx', ctypes.c_int32), ('ny', ctypes.c_int32),
                ('row_begin', ctypes.c_int32), ('row_count', ctypes.c_int32),
                ('batch', ctypes.c_int32), ('use_cds', ctypes.c_int32),
                ('dt', ctypes.c_double), ('dx', ctypes.c_double), ('dy', ctypes.c_double),
                ('rho', ctypes.c_double), ('mi', ctypes.c_double), ('bc', ctypes.c_double),
                ('dt_dev', ctypes.c_void_p), ('bc_dev', ctypes.c_void_p)]


class LdcSolverOptions(ctypes.Structure):
    _fields_ = [('snes_rtol', ctypes.c_double), ('snes_atol', ctypes.c_double),
                ('snes_stol', ctypes.c_double), ('snes_divtol', ctypes.c_double),
                ('snes_max_it', ctypes.c_int32), ('snes_max_funcs', ctypes.c_int32),
                ('ksp_rtol', ctypes.c_double), ('ksp_atol', ctypes.c_double),
                ('ksp_dtol', ctypes.c_double), ('ksp_stagnation', ctypes.c_double),
                ('ksp_max_it', ctypes.c_int32), ('gmres_restart', ctypes.c_int32),
                ('pc_type', ctypes.c_int32), ('pc_sweeps', ctypes.c_int32),
                ('mg_levels', ctypes.c_int32), ('mg_coarse_sweeps', ctypes.c_int32),
                ('mg_min_size', ctypes.c_int32), ('residual_variant', ctypes.c_int32),
                ('gmres_refine', ctypes.c_int32), ('use_cuda_graphs', ctypes.c_int32),
                ('pc_fp32', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('pc_omega_u', ctypes.c_double), ('pc_omega_p', ctypes.c_double)]


class LdcNewtonReport(ctypes.Structure):
    _fields_ = [('reason', ctypes.c_int32), ('iterations', ctypes.c_int32),
                ('function_evaluations', ctypes.c_int32), ('linear_iterations', ctypes.c_int32),
                ('ksp_reason', ctypes.c_int32), ('jacobian_evaluations', ctypes.c_int32),
                ('fnorm0', ctypes.c_double), ('fnorm', ctypes.c_double),
                ('xnorm', ctypes.c_double), ('ynorm', ctypes.c_double)]


class LdcPeerHalo(ctypes.Structure):
    _fields_ = [('prev_ghost', ctypes.c_void_p), ('next_ghost', ctypes.c_void_p),
                ('my_flags', ctypes.c_void_p), ('prev_flag', ctypes.c_void_p),
                ('next_flag', ctypes.c_void_p), ('epoch', ctypes.c_uint64),
                ('status_dev', ctypes.c_void_p)]


PC_NONE, PC_VANKA, PC_MG_VANKA = 0, 1, 2

# every symbol include/ldc_b200.h declares: name -> (restype, argtypes)
_vp = ctypes.c_void_p
_i32, _i64 = ctypes.c_int32, ctypes.c_int64
_grid_p = ctypes.POINTER(LdcGrid)
SYMBOLS = {
    'ldc_last_error': (ctypes.c_char_p, []),
    'ldc_version': (ctypes.c_int, []),
    'ldc_device_info': (ctypes.c_int, [c_int32_p, c_int32_p, c_int32_p, ctypes.POINTER(_i64)]),
    'ldc_residual_work_doubles': (_i64, [_grid_p]),
    'ldc_residual': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    'ldc_residual_host': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp, _i32]),
    'ldc_residual_peer_push': (ctypes.c_int, [_grid_p, ctypes.POINTER(LdcPeerHalo), _vp, _vp]),
    'ldc_residual_peer': (ctypes.c_int, [_grid_p, ctypes.POINTER(LdcPeerHalo), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ldc_peer_ctx_create': (ctypes.c_int, [_i32, ctypes.POINTER(_vp)]),
    'ldc_peer_ctx_destroy': (None, [_vp]),
    'ldc_peer_ctx_join': (ctypes.c_int, [_vp, _vp]),
    'ldc_residual_peer_step': (ctypes.c_int, [_vp, _grid_p, ctypes.POINTER(LdcPeerHalo), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ldc_pack_X': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp, _vp]),
    'ldc_unpack_X': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp, _vp]),
    'ldc_mask_nnz': (_i64, [_i32, _i32, _i32]),
    'ldc_mask_csr': (ctypes.c_int, [_i32, _i32, _i32, _vp, _vp, _vp]),
    'ldc_fd_jacobian': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp, _vp]),
    'ldc_fd_jacobian_compact': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp, _vp]),
    'ldc_spmv_compact': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp]),
    'ldc_spmv_mask': (ctypes.c_int, [_grid_p, _vp, _vp, _vp, _vp]),
    'ldc_spmv_csr': (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ldc_solver_default_options': (None, [ctypes.POINTER(LdcSolverOptions)]),
    'ldc_solver_create': (ctypes.c_int, [_grid_p, ctypes.POINTER(LdcSolverOptions), ctypes.POINTER(_vp)]),
    'ldc_solver_destroy': (None, [_vp]),
    'ldc_solver_state_dev': (_vp, [_vp]),
    'ldc_solver_set_dt': (ctypes.c_int, [_vp, c_double_p]),
    'ldc_solver_set_bc': (ctypes.c_int, [_vp, c_double_p]),
    'ldc_solver_set_state': (ctypes.c_int, [_vp, _vp, _vp]),
    'ldc_solver_get_state': (ctypes.c_int, [_vp, _vp, _vp]),
    'ldc_solver_set_comm': (ctypes.c_int, [_vp, _vp]),
    'ldc_solver_num_levels': (ctypes.c_int, [_vp]),
    'ldc_solver_set_keep_diverged': (ctypes.c_int, [_vp, _i32]),
    'ldc_solver_plan_levels': (ctypes.c_int, [_grid_p, ctypes.POINTER(LdcSolverOptions), c_int32_p, c_int32_p]),
    'ldc_solver_counters': (ctypes.c_int, [_vp, ctypes.POINTER(_i64)]),
    'ldc_solver_time_ops': (ctypes.c_int, [_vp, _i32, c_double_p, _vp]),
    'ldc_solver_linear_solve': (ctypes.c_int, [_vp, _vp, _vp, c_int32_p, c_int32_p, c_double_p, _vp]),
    'ldc_solver_step': (ctypes.c_int, [_vp, c_int32_p, ctypes.POINTER(LdcNewtonReport), c_double_p, _vp]),
}

_lib = None


def load_library():
    """Load libldc_b200.so and bind every declared symbol.  Raises RuntimeError (same wording
    as the reference's native forwarders, residual_function/cpp_residual_function.py:1-8)
    when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("CUDA module is not compiled. (%s missing; run "
                           "`python lid_driven_cavity_problem_b200/csrc/build.py`)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in list(SYMBOLS.items()):
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def lib():
    return load_library()


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("the cuda backend needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def check(rc, what=''):
    if rc != 0:
        msg = lib().ldc_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what or 'libldc_b200 call', rc,
                                                 msg.decode() if msg else '?'))


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def grid_from_graph(graph, batch=1, row_begin=0, row_count=None, dt_dev=None, bc_dev=None):
    nx, ny = int(graph.pressure_mesh.nx), int(graph.pressure_mesh.ny)
    g = LdcGrid()
    g.nx, g.ny = nx, ny
    g.row_begin = row_begin
    g.row_count = ny if row_count is None else row_count
    g.batch = batch
    g.use_cds = int(bool(getattr(graph, 'use_cds', False)))
    g.dt, g.dx, g.dy = float(graph.dt), float(graph.dx), float(graph.dy)
    g.rho, g.mi, g.bc = float(graph.rho), float(graph.mi), float(graph.bc)
    g.dt_dev = dt_dev
    g.bc_dev = bc_dev
    return g


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None
