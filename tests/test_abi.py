"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/ldc_b200.h declares (no compute calls: there is no GPU in the build container)."""
import os
import re

import pytest

from conftest import REPO


def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'ldc_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ldc_[A-Za-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), "libldc_b200.so does not export %s" % name
    bound = set(_native.SYMBOLS)
    assert bound == set(declared), (bound ^ set(declared))
    assert lib.ldc_version() >= 100


def test_mask_nnz_closed_form_host_side():
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.load_library()
    assert lib.ldc_mask_nnz(1024, 1024, 3) == 66023424
    assert lib.ldc_mask_nnz(4096, 4096, 3) == 1056817152
    assert lib.ldc_mask_nnz(5, 5, 1) == 155
    assert lib.ldc_mask_nnz(2, 4, 3) == -1      # reference: "Crashing" TODO


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    import numpy as np
    from conftest import synthetic_case
    from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function as crf
    g, X = synthetic_case(8, 8)
    with pytest.raises(RuntimeError):
        crf.residual_function(X, g)


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, 'lid_driven_cavity_problem_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                text = open(os.path.join(root, f)).read()
                assert 'ldc_oracle' not in text and 'from oracle' not in text, os.path.join(root, f)


def test_option_struct_layout_and_defaults():
    """ctypes mirror of ldc_solver_options: same size as the C struct (checked through the defaults
    the library writes) and the documented defaults"""
    import ctypes
    from lid_driven_cavity_problem_b200 import _native
    o = _native.LdcSolverOptions()
    _native.load_library().ldc_solver_default_options(ctypes.byref(o))
    assert (o.snes_rtol, o.snes_atol, o.snes_stol, o.snes_max_it) == (1e-4, 1e-4, 1e-4, 50)   # petsc_solver_wrapper.py:135
    assert o.gmres_restart == 30 and o.pc_type == _native.PC_MG_VANKA
    assert o.pc_omega_u == 0.8 and o.pc_omega_p == 0.8      # last fields: a layout slip would garble them
    assert ctypes.sizeof(_native.LdcGrid) == 88 and ctypes.sizeof(_native.LdcNewtonReport) == 56
    assert ctypes.sizeof(_native.LdcPeerHalo) == 56


def test_host_staging_copy_is_exact_on_both_paths():
    """the plug-in's host staging copy (torch threads, or the own pool when torchrun pinned torch to
    one thread) moves every element, including a ragged tail"""
    import numpy as np
    import torch
    from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function as crf
    before = torch.get_num_threads()
    try:
        for nt in (max(before, 2), 1):
            torch.set_num_threads(nt)
            for n in (1, 1000, (1 << 18) + 5, 3 * 700 * 700 + 1):
                src = np.random.default_rng(n).random(n)
                dst = np.full(n, -1.0)
                crf._parallel_copy(dst, src, torch)
                assert np.array_equal(dst, src), (nt, n)
    finally:
        torch.set_num_threads(before)


def test_argument_validation_happens_before_any_device_work():
    """invalid calls are refused with LDC_ERR_INVALID and a message before CUDA is touched, so this
    runs without a device; and with no device a valid call fails loudly instead of computing on the host"""
    import ctypes
    import numpy as np
    import torch
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.load_library()
    g = _native.LdcGrid()
    g.nx, g.ny, g.row_begin, g.row_count, g.batch = 8, 8, 0, 8, 1
    g.dt, g.dx, g.dy, g.rho, g.mi, g.bc = 1e-2, 1 / 8, 1 / 8, 1.0, 1.0, 10.0
    X, F = np.zeros(3 * 64), np.zeros(3 * 64)
    px, pf = X.ctypes.data_as(ctypes.c_void_p), F.ctypes.data_as(ctypes.c_void_p)

    def refused(rc, text):
        return rc == -1 and text in lib.ldc_last_error()
    assert refused(lib.ldc_residual(g, None, px, px, pf, None, None, 0, None), b'null device pointer')
    assert refused(lib.ldc_residual(g, px, px, px, px, None, None, 0, None), b'must not alias')
    assert refused(lib.ldc_residual(g, px, px, px, pf, None, None, 7, None), b'unknown variant')
    assert refused(lib.ldc_residual(g, px, px, px, pf, pf, None, 0, None), b'without work buffer')
    g.nx = 2
    assert refused(lib.ldc_residual(g, px, px, px, pf, None, None, 0, None), b'too small')
    g.nx, g.row_count = 8, 9
    assert refused(lib.ldc_residual(g, px, px, px, pf, None, None, 0, None), b'outside grid')
    g.row_count = 8
    ph = _native.LdcPeerHalo()
    assert refused(lib.ldc_residual_peer(g, ph, px, px, px, pf, None, None, None), b'epoch starts at 1')
    ph.epoch = 1
    ph.next_ghost = px        # a full grid has no neighbour strip
    assert refused(lib.ldc_residual_peer(g, ph, px, px, px, pf, None, None, None), b'neighbour strip exists')
    g.batch = 2
    assert refused(lib.ldc_residual_peer(g, ph, px, px, px, pf, None, None, None), b'batch must be 1')
    assert refused(lib.ldc_residual_peer_push(g, ph, px, None), b'batch must be 1')
    g.batch = 1
    assert refused(lib.ldc_residual_peer_push(g, ph, px, None), b'neighbour strip exists')
    assert lib.ldc_residual_work_doubles(g) > 0
    # the host-buffer entry point and the helpers added in round 2
    assert refused(lib.ldc_residual_host(g, None, px, px, pf, 0), b'null host pointer')
    assert refused(lib.ldc_residual_host(g, px, px, px, px, 0), b'must not alias')
    g.row_begin, g.row_count = 2, 4
    assert refused(lib.ldc_residual_host(g, px, px, px, pf, 0), b'needs the full grid')
    g.row_begin, g.row_count = 0, 8
    g.batch = 2
    assert refused(lib.ldc_residual_host(g, px, px, px, pf, 0), b'one cavity per call')
    g.batch = 1
    assert refused(lib.ldc_peer_ctx_create(0, ctypes.byref(ctypes.c_void_p())), b'lag must be in')
    assert refused(lib.ldc_solver_plan_levels(g, None, None, None), b'null argument')
    assert refused(lib.ldc_solver_time_ops(None, 1, None, None), b'bad argument')
    assert refused(lib.ldc_solver_set_keep_diverged(None, 1), b'null solver')
    if not torch.cuda.is_available():
        rc = lib.ldc_residual_host(g, px, px, px, pf, 0)
        assert rc == -2 and b'CUDA error' in lib.ldc_last_error()
        assert not F.any()
        rc = lib.ldc_residual(g, px, px, px, pf, None, None, 0, None)
        assert rc == -2 and b'CUDA error' in lib.ldc_last_error()
        assert not F.any()                  # nothing was computed anywhere
