"""Per-kernel timing sweep on one GPU (CUDA events, L2-cold by rotating buffer sets).

    python tools/kernel_sweep.py [--sizes 256,1024,4096] [--variants 1,2]

Prints one JSON line per (kernel, size, variant): ms, Mcells/s, algorithmic GB/s and the
fraction of the measured HBM peak (MEASURED_PEAKS.json).  This is the equivalent of the
reference's experiments/benchmarks.py sweep for the individual device kernels.
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))


def hbm_peak():
    try:
        return json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs'], 'measured'
    except Exception:
        return 6650.0, 'fallback'


def time_cuda(fn, n_sets, iters, warmup=3):
    import torch
    for k in range(warmup):
        fn(k % n_sets)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(iters):
        fn(k % n_sets)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='100,256,1024,2048,4096')
    ap.add_argument('--variants', default='1,2')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--kernels', default='residual,fd,spmv')
    args = ap.parse_args()
    import torch
    from conftest import synthetic_case
    from lid_driven_cavity_problem_b200 import _native
    lib = _native.lib()
    peak, peak_kind = hbm_peak()
    for n in [int(s) for s in args.sizes.split(',')]:
        g, X = synthetic_case(n, n)
        N = n * n
        res_bytes = 8 * (8 * N - 2 * n)
        # enough rotating buffer sets to exceed the 126 MB L2 several times over
        n_sets = max(2, min(16, int(600e6 // res_bytes) + 1))
        Xs = [torch.from_numpy(X).cuda() + float(k) for k in range(n_sets)]
        Fs = [torch.empty_like(Xs[0]) for _ in range(n_sets)]
        Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
        Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
        Uds = [Ud.clone() for _ in range(n_sets)]
        Vds = [Vd.clone() for _ in range(n_sets)]
        grid = _native.grid_from_graph(g)
        st = _native.current_stream()
        if 'residual' in args.kernels:
            for v in [int(s) for s in args.variants.split(',')]:
                def run(k):
                    lib.ldc_residual(grid, _native.ptr(Xs[k]), _native.ptr(Uds[k]), _native.ptr(Vds[k]),
                                     _native.ptr(Fs[k]), None, None, v, st)
                ms = time_cuda(run, n_sets, args.iters)
                gbs = res_bytes / ms / 1e6
                print(json.dumps(dict(kernel='residual', variant=v, n=n, ms=round(ms, 5),
                                      mcells_s=round(N / ms / 1e3, 1), gbs=round(gbs, 1),
                                      frac=round(gbs / peak, 4), peak=peak_kind, sets=n_sets)), flush=True)
        nnz = int(lib.ldc_mask_nnz(n, n, 3))
        if nnz * 8 * 2 > 60e9:
            continue
        j_sets = 2 if nnz * 8 > 300e6 else max(2, min(8, int(600e6 // (nnz * 8)) + 1))
        Js = [torch.empty(nnz, dtype=torch.float64, device='cuda') for _ in range(j_sets)]
        if 'fd' in args.kernels or 'spmv' in args.kernels:
            def run_fd(k):
                lib.ldc_fd_jacobian(grid, _native.ptr(Xs[k % n_sets]), _native.ptr(Uds[k % n_sets]),
                                    _native.ptr(Vds[k % n_sets]), _native.ptr(Js[k % j_sets]), st)
            ms = time_cuda(run_fd, j_sets, max(4, args.iters // 2))
            fd_bytes = res_bytes - 24 * N + 8 * nnz
            print(json.dumps(dict(kernel='fd_jacobian', n=n, ms=round(ms, 5), mcells_s=round(N / ms / 1e3, 1),
                                  gbs=round(fd_bytes / ms / 1e6, 1), frac=round(fd_bytes / ms / 1e6 / peak, 4),
                                  peak=peak_kind)), flush=True)
        if 'spmv' in args.kernels:
            xs = [torch.randn(3 * N, dtype=torch.float64, device='cuda') for _ in range(j_sets)]
            ys = [torch.empty_like(xs[0]) for _ in range(j_sets)]
            def run_spmv(k):
                lib.ldc_spmv_mask(grid, _native.ptr(Js[k]), _native.ptr(xs[k]), _native.ptr(ys[k]), st)
            ms = time_cuda(run_spmv, j_sets, args.iters)
            own = 8 * nnz + 48 * N
            canon = 12 * nnz + 4 * (3 * N + 1) + 16 * 3 * N
            print(json.dumps(dict(kernel='spmv_mask', n=n, ms=round(ms, 5), gbs_own=round(own / ms / 1e6, 1),
                                  frac_own=round(own / ms / 1e6 / peak, 4),
                                  gbs_canonical=round(canon / ms / 1e6, 1),
                                  frac_canonical=round(canon / ms / 1e6 / peak, 4), peak=peak_kind)), flush=True)
            if nnz * 4 < 8e9:
                from lid_driven_cavity_problem_b200.nonlinear_solver import _common
                indptr, indices = _common.jacobian_mask_device(n, n, 3)
                def run_csr(k):
                    lib.ldc_spmv_csr(3 * N, _native.ptr(indptr), _native.ptr(indices), _native.ptr(Js[k]),
                                     _native.ptr(xs[k]), _native.ptr(ys[k]), st)
                ms = time_cuda(run_csr, j_sets, args.iters)
                print(json.dumps(dict(kernel='spmv_csr', n=n, ms=round(ms, 5),
                                      gbs_canonical=round(canon / ms / 1e6, 1),
                                      frac_canonical=round(canon / ms / 1e6 / peak, 4), peak=peak_kind)), flush=True)
                del indptr, indices
        del Js, Xs, Fs
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
