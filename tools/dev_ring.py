"""Developer check of the residual kernel variants on a B200: parity of every variant against
variant 2 / the oracle on awkward shapes, then kernel-only timing (CUDA events, L2-cold buffer
rotation) for the tuning knobs of variant 4 (one subprocess per knob set, they are read once).

    python tools/dev_ring.py check            # parity sweep
    python tools/dev_ring.py time [nx ny batch variant]   # one timing line for the current env
    python tools/dev_ring.py sweep            # spawn `time` under several LDC_RING_* settings
"""
import itertools
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))


def _setup():
    import torch
    from lid_driven_cavity_problem_b200 import _native
    return torch, _native, _native.lib()


def run_variant(torch, native, lib, g, X, U, V, variant, batch=1, row_begin=0, row_count=None, x_off_rows=0,
                norm=False):
    nx = int(g.pressure_mesh.nx)
    grid = native.grid_from_graph(g, batch=batch, row_begin=row_begin, row_count=row_count)
    rows = grid.row_count
    F = torch.full((3 * nx * rows * batch,), float('nan'), dtype=torch.float64, device='cuda')
    n2 = work = None
    if norm:
        n2 = torch.zeros(batch, dtype=torch.float64, device='cuda')
        work = torch.empty(int(lib.ldc_residual_work_doubles(grid)), dtype=torch.float64, device='cuda')
    native.check(lib.ldc_residual(grid, X.data_ptr() + 8 * 3 * nx * x_off_rows, native.ptr(U), native.ptr(V),
                                  native.ptr(F), native.ptr(n2), native.ptr(work), variant,
                                  native.current_stream()), 'ldc_residual')
    return F, n2


def check():
    from conftest import synthetic_case
    from oracle import ldc_oracle as orc
    torch, native, lib = _setup()
    bad = 0
    shapes = [(4, 2), (4, 3), (6, 5), (8, 8), (62, 7), (64, 9), (66, 5), (124, 9), (126, 40), (128, 3), (130, 2),
              (100, 100), (250, 33), (256, 256), (1024, 1024), (2048, 64), (1022, 77), (64, 64), (4096, 512)]
    if os.environ.get('LDC_CHECK_SHAPES'):
        shapes = [tuple(int(v) for v in t.split('x')) for t in os.environ['LDC_CHECK_SHAPES'].split(',')]
    for nx, ny in shapes:
        for cds in (False, True):
            g, X = synthetic_case(nx, ny)
            g.use_cds = cds
            # dummy slots carry non-zero values in the middle of a Newton solve
            X[1::3][nx - 1::nx] = 0.3
            X[2::3][nx * (ny - 1):] = -0.2
            Xd = torch.from_numpy(X).cuda()
            Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
            Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
            verbose = bool(os.environ.get('LDC_CHECK_VERBOSE'))
            if verbose:
                torch.cuda.synchronize()
                print('  X %x U %x V %x' % (Xd.data_ptr(), Ud.data_ptr(), Vd.data_ptr()), flush=True)
                print(torch.cuda.memory_summary(abbreviated=True)[:0], end='')
                for seg in torch.cuda.memory_snapshot():
                    print('   seg %x size %d' % (seg['address'], seg['total_size']), flush=True)
            F2, _ = run_variant(torch, native, lib, g, Xd, Ud, Vd, 2)
            if verbose:
                torch.cuda.synchronize()
                print('  v2 done F %x' % F2.data_ptr(), flush=True)
            F4, n2 = run_variant(torch, native, lib, g, Xd, Ud, Vd, 4, norm=True)
            if verbose:
                torch.cuda.synchronize()
                print('  v4 done F %x' % F4.data_ptr(), flush=True)
            ok = torch.equal(F2, F4)
            if nx * ny <= 300000:
                Fo = orc.residual_function(X, g, nthreads=0)
                ok = ok and np.array_equal(F4.cpu().numpy(), Fo)
            ref = float((F2 * F2).sum().item())
            okn = abs(n2.item() - ref) <= 1e-12 * ref
            print('full', nx, ny, 'cds' if cds else 'uds', 'OK' if ok and okn else 'MISMATCH', flush=True)
            if not (ok and okn):
                bad += 1
                d = (F2 != F4).nonzero().flatten()
                print('   first diffs at', d[:12].tolist(), 'count', d.numel(), 'norm', n2.item(), ref)
    # non power-of-two spacing and rho, mi != 1 (generic arithmetic path)
    for nx, ny, rho, mi in [(100, 60, 1.3, 0.7), (64, 64, 1.0, 0.01), (256, 128, 2.0, 1.0)]:
        g, X = synthetic_case(nx, ny)
        g.rho, g.mi = rho, mi
        Xd = torch.from_numpy(X).cuda()
        Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
        Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
        F4, _ = run_variant(torch, native, lib, g, Xd, Ud, Vd, 4)
        Fo = orc.residual_function(X, g, nthreads=0)
        ok = np.array_equal(F4.cpu().numpy(), Fo)
        print('phys', nx, ny, rho, mi, 'OK' if ok else 'MISMATCH', flush=True)
        bad += 0 if ok else 1
    # batch (odd row count -> odd per-instance U_old stride, tail over-read guard)
    for nx, ny, B in [(64, 64, 7), (64, 33, 5), (30, 9, 3), (126, 5, 4)]:
        Xs, Us, Vs = [], [], []
        bcs = np.linspace(100.0, 5000.0, B)
        for b in range(B):
            g, X = synthetic_case(nx, ny, bc=float(bcs[b]), seed=b)
            Xs.append(X); Us.append(g.ns_x_mesh.phi_old); Vs.append(g.ns_y_mesh.phi_old)
        g, _ = synthetic_case(nx, ny)
        Xd = torch.from_numpy(np.concatenate(Xs)).cuda()
        Ud = torch.from_numpy(np.concatenate(Us)).cuda()
        Vd = torch.from_numpy(np.concatenate(Vs)).cuda()
        bc_dev = torch.from_numpy(bcs).cuda()
        outs = []
        for v in (2, 4):
            grid = native.grid_from_graph(g, batch=B)
            grid.bc_dev = bc_dev.data_ptr()
            F = torch.full_like(Xd, float('nan'))
            n2 = torch.zeros(B, dtype=torch.float64, device='cuda')
            work = torch.empty(int(lib.ldc_residual_work_doubles(grid)), dtype=torch.float64, device='cuda')
            native.check(lib.ldc_residual(grid, native.ptr(Xd), native.ptr(Ud), native.ptr(Vd), native.ptr(F),
                                          native.ptr(n2), native.ptr(work), v, native.current_stream()))
            outs.append((F, n2))
        ok = torch.equal(outs[0][0], outs[1][0])
        okn = bool((abs(outs[0][1] - outs[1][1]) <= 1e-12 * outs[0][1]).all())
        print('batch', nx, ny, B, 'OK' if ok and okn else 'MISMATCH', flush=True)
        bad += 0 if (ok and okn) else 1
    # row strips with ghost rows
    for nx, ny, bounds in [(96, 70, [0, 17, 18, 40, 70]), (1024, 64, [0, 1, 33, 64]), (64, 50, [0, 13, 14, 32, 50])]:
        g, X = synthetic_case(nx, ny)
        Xd = torch.from_numpy(X).cuda()
        Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
        Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
        Ffull, _ = run_variant(torch, native, lib, g, Xd, Ud, Vd, 2)
        U_old = g.ns_x_mesh.phi_old.reshape(ny, nx - 1)
        V_old = g.ns_y_mesh.phi_old.reshape(ny - 1, nx)
        ok = True
        for j0, j1 in zip(bounds[:-1], bounds[1:]):
            lo, hi = max(j0 - 1, 0), min(j1 + 1, ny)
            Xl = torch.from_numpy(X[3 * nx * lo:3 * nx * hi].copy()).cuda()
            Ul = torch.from_numpy(U_old[j0:j1].copy().ravel()).cuda()
            Vl = torch.from_numpy(np.r_[V_old[j0:min(j1, ny - 1)].ravel(), np.zeros(2)]).cuda()
            Fl, _ = run_variant(torch, native, lib, g, Xl, Ul, Vl, 4, row_begin=j0, row_count=j1 - j0,
                                x_off_rows=j0 - lo)
            ok = ok and torch.equal(Fl, Ffull[3 * nx * j0:3 * nx * j1])
        print('strips', nx, ny, 'OK' if ok else 'MISMATCH', flush=True)
        bad += 0 if ok else 1
    print('RESULT', 'all ok' if bad == 0 else '%d failures' % bad)
    return bad


def time_one(nx, ny, batch, variant, iters=None):
    from conftest import synthetic_case
    torch, native, lib = _setup()
    g, X = synthetic_case(nx, ny)
    N = nx * ny
    bytes_alg = 8 * (8 * N - nx - ny) * batch
    n_sets = max(2, min(12, int(700e6 // bytes_alg) + 1))
    Xd = torch.from_numpy(np.tile(X, batch)).cuda()
    Xs = [Xd + float(k) for k in range(n_sets)]
    Fs = [torch.empty_like(Xd) for _ in range(n_sets)]
    Us = [torch.from_numpy(np.tile(g.ns_x_mesh.phi_old, batch)).cuda() for _ in range(n_sets)]
    Vs = [torch.from_numpy(np.tile(g.ns_y_mesh.phi_old, batch)).cuda() for _ in range(n_sets)]
    grid = native.grid_from_graph(g, batch=batch)
    st = native.current_stream()

    def run(k):
        lib.ldc_residual(grid, native.ptr(Xs[k]), native.ptr(Us[k]), native.ptr(Vs[k]), native.ptr(Fs[k]),
                         None, None, variant, st)
    iters = iters or (100 if N * batch < 4e6 else 30)
    for k in range(10):
        run(k % n_sets)
    torch.cuda.synchronize()
    reps = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(iters):
            run(k % n_sets)
        e1.record()
        torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1) / iters)
    ms = float(np.median(reps))
    out = {"nx": nx, "ny": ny, "batch": batch, "variant": variant, "us": round(ms * 1e3, 3),
           "us_min": round(min(reps) * 1e3, 3), "frac": round(bytes_alg / ms / 1e6 / 6534.8, 4),
           "env": {k: v for k, v in os.environ.items() if k.startswith('LDC_')}}
    print(json.dumps(out), flush=True)


def sweep():
    cases = [(1024, 1024, 1), (4096, 4096, 1), (64, 64, 512), (256, 256, 1)]
    runs = []
    for nx, ny, b in cases:
        runs.append(({}, nx, ny, b, 2))
    for cfg in range(7):
        runs.append(({'LDC_RING_CFG': str(cfg)}, 1024, 1024, 1, 4))
    for cfg, rows in itertools.product((0, 4), (7, 8, 9, 10, 11, 12, 13, 14, 16)):
        runs.append(({'LDC_RING_CFG': str(cfg), 'LDC_RING_ROWS': str(rows)}, 1024, 1024, 1, 4))
    for cfg in (0, 1, 2, 4):
        runs.append(({'LDC_RING_CFG': str(cfg)}, 4096, 4096, 1, 4))
        runs.append(({'LDC_RING_CFG': str(cfg)}, 64, 64, 512, 4))
    runs.append(({'LDC_NO_PDL': '1'}, 1024, 1024, 1, 4))
    runs.append(({}, 256, 256, 1, 4))
    runs.append(({}, 100, 100, 1, 4))
    runs.append(({}, 2048, 2048, 1, 4))
    for env, nx, ny, b, v in runs:
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), 'time', str(nx), str(ny), str(b), str(v)], env=e)


def one(nx, ny, variant, cds, norm, reps):
    """one shape / variant in isolation, synchronised after every launch (to pin a fault)"""
    from conftest import synthetic_case
    torch, native, lib = _setup()
    g, X = synthetic_case(nx, ny)
    g.use_cds = bool(cds)
    Xd = torch.from_numpy(X).cuda()
    Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
    Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
    F2, _ = run_variant(torch, native, lib, g, Xd, Ud, Vd, 2)
    torch.cuda.synchronize()
    for r in range(reps):
        F, n2 = run_variant(torch, native, lib, g, Xd, Ud, Vd, variant, norm=bool(norm))
        torch.cuda.synchronize()
        if not torch.equal(F, F2):
            print('MISMATCH rep', r)
    print('one', nx, ny, variant, cds, norm, 'done', flush=True)


if __name__ == '__main__':
    cmd = sys.argv[1] if len(sys.argv) > 1 else 'check'
    if cmd == 'one':
        one(*[int(v) for v in sys.argv[2:8]])
        sys.exit(0)
    if cmd == 'check':
        sys.exit(1 if check() else 0)
    elif cmd == 'time':
        a = [int(v) for v in sys.argv[2:6]] if len(sys.argv) >= 6 else [1024, 1024, 1, 4]
        time_one(*a)
    elif cmd == 'sweep':
        sweep()
