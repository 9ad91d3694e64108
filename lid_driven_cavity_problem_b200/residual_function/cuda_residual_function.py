"""``residual_function(X, graph)`` on the B200 -- the `cuda` sibling of the reference's
numpy / numba / Cython / C++ / C++-OpenMP residual modules
(residual_function/cpp_residual_function.py:1-8 forwards to c++/residual_function.cpp:10).

Contract: pure (neither X nor graph is modified), returns a freshly allocated
length-3*nx*ny float64 vector.  ``X`` may be

* a numpy array / sequence (what PETSc's callback and scipy's fsolve pass,
  petsc_solver_wrapper.py:42-48, scipy_solver_wrapper.py:35): it is copied to the device
  by a pipelined library call (pinned staging, H2D, kernel per row block, D2H overlapped:
  ldc_residual_host, csrc/host_pipeline.cu), and a numpy array comes back;
* a CUDA ``torch.Tensor`` (float64): the result is a CUDA tensor and nothing crosses PCIe.

Reads exactly what the reference's native residual reads from ``graph``
(c++/residual_function.cpp:16-43): dt, dx, dy, rho, mi, bc, the mesh sizes and the
``phi_old`` arrays of the U and V meshes (+ ``use_cds``, numba_residual_function.py:247).
Unlike the C++ module, which reinterprets whatever buffer it is given as float64
(SURVEY App. G), X is validated and converted to C-contiguous float64.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from lid_driven_cavity_problem_b200 import _native

try:
    _native.load_library()
except RuntimeError:
    raise RuntimeError("CUDA module is not compiled.")

_staging = {}
_PINNED_RESULT_LIMIT = 512 << 20   # bytes


def _pinned(name, n, torch):
    key = (name, n)
    buf = _staging.get(key)
    if buf is None:
        if len(_staging) > 16:
            _staging.clear()
        buf = torch.empty(n, dtype=torch.float64, pin_memory=True)
        _staging[key] = buf
    return buf


def _as_f64(name, host, n):
    arr = np.ascontiguousarray(host, dtype=np.float64).reshape(-1)
    if arr.size != n:
        raise ValueError("%s has %d entries, expected %d" % (name, arr.size, n))
    return arr


_pool = None


def _copy_threads():
    ranks_here = int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1)
    return max(1, min(8, (os.cpu_count() or 1) // max(ranks_here, 1)))


def _parallel_copy(dst, src, torch):
    """dst[:] = src for flat contiguous float64 numpy arrays on several host threads.  torch's
    copy_ is the fast path (measured: 1024^2 round trip 2.2-2.4 ms with it, ~4 ms with the pool
    below), but its threading follows OMP_NUM_THREADS, which torchrun sets to 1 for every worker
    (the same round trip then takes 10 ms): in that case split the copy over an own small thread
    pool (numpy drops the GIL inside copyto)."""
    global _pool
    if torch.get_num_threads() > 1:
        torch.from_numpy(dst).copy_(torch.from_numpy(src))
        return
    n, nt = dst.size, _copy_threads()
    if nt == 1 or n < (1 << 18):
        np.copyto(dst, src)
        return
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix='ldc-stage')
    step = -(-n // nt)
    list(_pool.map(lambda a: np.copyto(dst[a:a + step], src[a:a + step]), range(0, n, step)))


def _to_device(name, host, n, torch):
    """host array -> device tensor: multi-threaded copy into a cached pinned buffer, then an
    asynchronous DMA on the current stream (the next array is staged while this one travels)"""
    arr = _as_f64(name, host, n)
    pin = _pinned(name, n, torch)
    _parallel_copy(pin.numpy(), arr, torch)
    return pin.to('cuda', non_blocking=True)


def residual_function_device(X_dev, graph, U_old_dev=None, V_old_dev=None, out=None, variant=0,
                             norm2_out=None, batch=1):
    """Device-to-device evaluation; every argument already lives in HBM."""
    torch = _native.require_cuda()
    lib = _native.lib()
    nx, ny = int(graph.pressure_mesh.nx), int(graph.pressure_mesh.ny)
    n = 3 * nx * ny * batch
    if X_dev.dtype != torch.float64 or not X_dev.is_cuda or X_dev.numel() != n:
        raise ValueError("X must be a CUDA float64 tensor of %d entries" % n)
    X_dev = X_dev.contiguous()
    if U_old_dev is None:
        U_old_dev = _to_device('U_old', graph.ns_x_mesh.phi_old, (nx - 1) * ny * batch, torch)
    if V_old_dev is None:
        V_old_dev = _to_device('V_old', graph.ns_y_mesh.phi_old, nx * (ny - 1) * batch, torch)
    F = torch.empty_like(X_dev) if out is None else out
    g = _native.grid_from_graph(graph, batch=batch)
    work = None
    if norm2_out is not None:
        work = torch.empty(int(lib.ldc_residual_work_doubles(g)), dtype=torch.float64, device='cuda')
    rc = lib.ldc_residual(g, _native.ptr(X_dev), _native.ptr(U_old_dev), _native.ptr(V_old_dev),
                          _native.ptr(F), _native.ptr(norm2_out), _native.ptr(work), variant,
                          _native.current_stream())
    _native.check(rc, 'ldc_residual')
    return F


def residual_function(X, graph):
    torch = _native.require_cuda()
    if isinstance(X, torch.Tensor):
        return residual_function_device(X, graph)
    # host arrays in, host array out: ONE library call (ldc_residual_host) that pipelines staging,
    # H2D, the kernel per row block, D2H and copy-out -- see csrc/host_pipeline.cu
    nx, ny = int(graph.pressure_mesh.nx), int(graph.pressure_mesh.ny)
    n = 3 * nx * ny
    Xh = _as_f64('X', X, n)
    Uh = _as_f64('U_old', graph.ns_x_mesh.phi_old, (nx - 1) * ny)
    Vh = _as_f64('V_old', graph.ns_y_mesh.phi_old, nx * (ny - 1))
    # Freshly allocated for every call, like the reference backends.  LDC_PINNED_RESULT=1 takes the
    # array out of torch's cache of page-locked host memory instead, so that the D2H copies land in
    # the result itself (ldc_residual_host detects a page-locked F): 10 % faster for a caller that
    # drops each result before the next call (1.34 vs 1.48 ms per 1024^2 call), 20 % SLOWER for one
    # that still holds the previous result (bench.py's loop: 1.82 vs 1.53 ms) -- hence opt-in.
    if os.environ.get('LDC_PINNED_RESULT') and 8 * n <= _PINNED_RESULT_LIMIT:
        result = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    else:
        result = np.empty(n, dtype=np.float64)
    g = _native.grid_from_graph(graph)
    rc = _native.lib().ldc_residual_host(g, Xh.ctypes.data, Uh.ctypes.data, Vh.ctypes.data, result.ctypes.data, 0)
    _native.check(rc, 'ldc_residual_host')
    return result
