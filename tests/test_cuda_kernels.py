"""GPU parity tests: every CUDA entry point against the CPU oracle and the golden vectors.
All calls go through the C-ABI (ctypes -> libldc_b200.so)."""
import numpy as np
import pytest

from conftest import make_graph, synthetic_case

pytestmark = pytest.mark.gpu

# residual kernels behind ldc_residual: 1 cell-per-thread, 2 register-prefetch march, 3 TMA row ring
# per block, 4 warp-autonomous bulk-copy rings (the default; odd nx falls back to 2)
VARIANTS = (1, 2, 3, 4)


@pytest.fixture(scope='module')
def cuda():
    import torch
    from lid_driven_cavity_problem_b200.residual_function import cuda_residual_function as crf
    from lid_driven_cavity_problem_b200.nonlinear_solver import _common
    from lid_driven_cavity_problem_b200 import _native
    return dict(torch=torch, crf=crf, common=_common, native=_native)


def _residual_variants(cuda, X, g):
    torch, crf = cuda['torch'], cuda['crf']
    Xd = torch.from_numpy(X).cuda()
    return [crf.residual_function_device(Xd, g, variant=v).cpu().numpy() for v in VARIANTS]


def test_reference_3x3_known_answer(cuda, golden):
    d = golden('residual_3x3_reference_test.npz')
    from lid_driven_cavity_problem_b200.staggered_grid import Graph
    g = Graph(1.0, 1.0, 3, 3, 0.01, 1.0, 1.0, 100.0, initial_P=100.0, initial_U=0.001, initial_V=-0.01)
    F = cuda['crf'].residual_function(d['X'], g)
    assert isinstance(F, np.ndarray) and F.dtype == np.float64
    assert np.array_equal(F, d['F'])                 # bit-exact
    assert np.allclose(F, d['F'], rtol=1e-6, atol=1e-4)  # the reference test's own tolerance
    for Fv in _residual_variants(cuda, d['X'], g):
        assert np.array_equal(Fv, d['F'])


def test_golden_random_cases_bit_exact(cuda, golden):
    d = golden('residual_random_cases.npz')
    for k in range(int(d['count'])):
        g = make_graph(d['params_%d' % k], 0.9, d['U_%d' % k], d['V_%d' % k], d['P_%d' % k])
        for tag in ('', '_dummy'):
            X, F = d['X%s_%d' % (tag, k)], d['F%s_%d' % (tag, k)]
            for v, Fv in zip(VARIANTS, _residual_variants(cuda, X, g)):
                assert np.array_equal(Fv, F), (k, tag, v)
        # packing
        common = cuda['common']
        assert np.array_equal(common._create_X(d['U_%d' % k], d['V_%d' % k], d['P_%d' % k], g), d['X_%d' % k])
        U, V, P = common._recover_X(d['X_dummy_%d' % k], g)
        assert np.array_equal(U, d['U_%d' % k]) and np.array_equal(V, d['V_%d' % k])
        assert np.array_equal(P, d['P_%d' % k])


@pytest.mark.parametrize("nx,ny", [(100, 100), (256, 256), (257, 131), (1024, 1024), (61, 1030), (124, 9),
                                   (126, 40), (250, 33), (4, 3), (2048, 64), (4096, 4096), (4, 2), (6, 5),
                                   (62, 7), (64, 9), (66, 5), (128, 3), (130, 2), (1022, 77), (4096, 512)])
def test_residual_vs_oracle(cuda, nx, ny):
    from oracle import ldc_oracle as orc
    g, X = synthetic_case(nx, ny)
    # dummy slots carry non-zero values in the middle of a Newton solve (rows F = X)
    X[1::3][nx - 1::nx] = 0.3
    X[2::3][nx * (ny - 1):] = -0.2
    F_ref = orc.residual_function(X, g, nthreads=0)
    for v, Fv in zip(VARIANTS, _residual_variants(cuda, X, g)):
        # north_star: 1e-12 relative (normwise); the kernels are in fact bit-exact
        assert np.abs(Fv - F_ref).max() <= 1e-12 * np.abs(F_ref).max(), (nx, ny, v)
        assert np.array_equal(Fv, F_ref), (nx, ny, v)
    g.use_cds = True
    F_cds = orc.residual_function(X, g, nthreads=0)
    for Fv in _residual_variants(cuda, X, g):
        assert np.array_equal(Fv, F_cds)


def test_residual_fused_norm(cuda):
    torch, crf = cuda['torch'], cuda['crf']
    for nx, ny in [(100, 100), (513, 300)]:
        g, X = synthetic_case(nx, ny)
        Xd = torch.from_numpy(X).cuda()
        for v in VARIANTS:
            n2 = torch.zeros(1, dtype=torch.float64, device='cuda')
            F = crf.residual_function_device(Xd, g, variant=v, norm2_out=n2)
            ref = float(np.dot(F.cpu().numpy(), F.cpu().numpy()))
            assert abs(n2.item() - ref) <= 1e-12 * ref


def test_residual_batch_ensemble(cuda):
    """Reynolds ensemble: instances back to back, per-instance lid speed on the device."""
    from oracle import ldc_oracle as orc
    torch, native = cuda['torch'], cuda['native']
    nx = ny = 64
    B = 5
    bcs = np.array([100.0, 400.0, 1000.0, 3200.0, 10000.0])
    Xs, Us, Vs, Fs = [], [], [], []
    for b in range(B):
        g, X = synthetic_case(nx, ny, bc=bcs[b], seed=b)
        Xs.append(X); Us.append(g.ns_x_mesh.phi_old); Vs.append(g.ns_y_mesh.phi_old)
        Fs.append(orc.residual_function(X, g))
    g, _ = synthetic_case(nx, ny)
    grid = native.grid_from_graph(g, batch=B)
    bc_dev = torch.from_numpy(bcs).cuda()
    grid.bc_dev = bc_dev.data_ptr()
    Xd = torch.from_numpy(np.concatenate(Xs)).cuda()
    Ud = torch.from_numpy(np.concatenate(Us)).cuda()
    Vd = torch.from_numpy(np.concatenate(Vs)).cuda()
    for v in VARIANTS:
        F = torch.empty_like(Xd)
        native.check(native.lib().ldc_residual(grid, native.ptr(Xd), native.ptr(Ud), native.ptr(Vd),
                                               native.ptr(F), None, None, v, native.current_stream()))
        assert np.array_equal(F.cpu().numpy(), np.concatenate(Fs))


def test_residual_row_strips(cuda):
    """Row-strip decomposition: each strip with its ghost rows reproduces its slice."""
    from oracle import ldc_oracle as orc
    torch, native = cuda['torch'], cuda['native']
    nx, ny = 96, 70
    g, X = synthetic_case(nx, ny)
    F_ref = orc.residual_function(X, g)
    U_old = g.ns_x_mesh.phi_old.reshape(ny, nx - 1)
    V_old = g.ns_y_mesh.phi_old.reshape(ny - 1, nx)
    bounds = [0, 17, 18, 40, 70]
    for v in VARIANTS:
        out = np.empty_like(F_ref)
        for j0, j1 in zip(bounds[:-1], bounds[1:]):
            lo, hi = max(j0 - 1, 0), min(j1 + 1, ny)
            Xl = torch.from_numpy(X[3 * nx * lo:3 * nx * hi].copy()).cuda()
            Ul = torch.from_numpy(U_old[j0:j1].copy().ravel()).cuda()
            Vl = torch.from_numpy(V_old[j0:min(j1, ny - 1)].copy().ravel()).cuda()
            if Vl.numel() == 0:
                Vl = torch.zeros(1, dtype=torch.float64, device='cuda')
            Fl = torch.full((3 * nx * (j1 - j0),), np.nan, dtype=torch.float64, device='cuda')
            grid = native.grid_from_graph(g, row_begin=j0, row_count=j1 - j0)
            x_ptr = Xl.data_ptr() + 8 * 3 * nx * (j0 - lo)
            native.check(native.lib().ldc_residual(grid, x_ptr, native.ptr(Ul), native.ptr(Vl), native.ptr(Fl),
                                                   None, None, v, native.current_stream()))
            out[3 * nx * j0:3 * nx * j1] = Fl.cpu().numpy()
        assert np.array_equal(out, F_ref), v


@pytest.mark.parametrize("nx,ny,bounds", [(96, 70, [0, 17, 18, 40, 70]), (250, 128, [0, 64, 128]),
                                          (1024, 512, [0, 128, 256, 384, 512]), (64, 12, [0, 1, 2, 5, 12]),
                                          (257, 40, [0, 20, 40])])
def test_residual_peer_halo_strips_one_gpu(cuda, nx, ny, bounds):
    """ldc_residual_peer_push + ldc_residual_peer: every strip pushes its boundary rows into the
    neighbours' ghost rows and publishes the evaluation number, the residual kernels wait for the
    words and read the ghosts locally.  All strips live on one GPU here, so all pushes of an
    evaluation are launched (and complete, stream order) before the first residual kernel: kernels
    that wait for each other must not be separate launches on one device.  The overlapped form
    runs in `bench.py --gpus N`, which compares every strip with a single-GPU evaluation.  Result
    bit-exact, twice in a row (second epoch on changed X), status word clean.  Odd nx takes the
    register-prefetch kernel, even nx the ring kernel."""
    from oracle import ldc_oracle as orc
    from lid_driven_cavity_problem_b200 import peer_halo
    torch, native = cuda['torch'], cuda['native']
    g, X = synthetic_case(nx, ny)
    U_old = g.ns_x_mesh.phi_old.reshape(ny, nx - 1)
    V_old = g.ns_y_mesh.phi_old.reshape(ny - 1, nx)
    spans = list(zip(bounds[:-1], bounds[1:]))
    rows_max = max(j1 - j0 for j0, j1 in spans)
    nbytes = peer_halo.region_bytes(nx, rows_max, 1)
    regions = [torch.zeros(nbytes // 8, dtype=torch.float64, device='cuda') for _ in spans]
    strips, Us, Vs, Fs = [], [], [], []
    for r, (j0, j1) in enumerate(spans):
        grid = native.grid_from_graph(g, row_begin=j0, row_count=j1 - j0)
        prev = regions[r - 1].data_ptr() if r > 0 else 0
        nxt = regions[r + 1].data_ptr() if r + 1 < len(spans) else 0
        prev_rows = spans[r - 1][1] - spans[r - 1][0] if r > 0 else 0
        strips.append(peer_halo.PeerStripResidual(grid, regions[r].data_ptr(), prev, nxt, prev_rows, rows_max))
        Us.append(torch.from_numpy(U_old[j0:j1].copy().ravel()).cuda())
        Vl = V_old[j0:min(j1, ny - 1)].copy().ravel()
        Vs.append(torch.from_numpy(Vl if Vl.size else np.zeros(2)).cuda())
        Fs.append(torch.full((3 * nx * (j1 - j0),), np.nan, dtype=torch.float64, device='cuda'))
    for epoch in range(2):
        Xe = X * (1.0 + 0.25 * epoch)
        F_ref = orc.residual_function(Xe, g)
        for r, (j0, j1) in enumerate(spans):
            strips[r].x_tensor(0).copy_(torch.from_numpy(Xe[3 * nx * j0:3 * nx * j1].copy()))
        torch.cuda.synchronize()
        n2 = [torch.zeros(1, dtype=torch.float64, device='cuda') for _ in spans]
        work = [torch.zeros(int(native.lib().ldc_residual_work_doubles(s.grid)), dtype=torch.float64,
                            device='cuda') for s in strips]
        launches = [strips[r].bind(0, Us[r], Vs[r], Fs[r], n2[r], work[r]) for r in range(len(spans))]
        st = native.current_stream()
        for launch in launches:
            launch(st, kernel=False)                 # all pushes first ...
        for launch in launches:
            launch(st, push=False)                   # ... then the residual kernels
        torch.cuda.synchronize()
        out = np.concatenate([F.cpu().numpy() for F in Fs])
        assert not any(s.status() for s in strips)
        assert np.array_equal(out, F_ref), epoch
        total = sum(float(t.item()) for t in n2)
        assert abs(total - float(np.dot(F_ref, F_ref))) <= 1e-12 * float(np.dot(F_ref, F_ref))


def test_residual_peer_halo_times_out_instead_of_hanging(cuda):
    """a strip whose neighbour never shows up finishes (status word set) instead of hanging"""
    from lid_driven_cavity_problem_b200 import peer_halo
    torch, native = cuda['torch'], cuda['native']
    nx, ny = 64, 32
    g, X = synthetic_case(nx, ny)
    nbytes = peer_halo.region_bytes(nx, 16, 1)
    mine = torch.zeros(nbytes // 8, dtype=torch.float64, device='cuda')
    ghost_town = torch.zeros(nbytes // 8, dtype=torch.float64, device='cuda')
    grid = native.grid_from_graph(g, row_begin=0, row_count=16)
    strip = peer_halo.PeerStripResidual(grid, mine.data_ptr(), 0, ghost_town.data_ptr(), 0, 16)
    strip.x_tensor(0).copy_(torch.from_numpy(X[:3 * nx * 16].copy()))
    U = torch.from_numpy(g.ns_x_mesh.phi_old[:(nx - 1) * 16].copy()).cuda()
    V = torch.from_numpy(g.ns_y_mesh.phi_old[:nx * 16].copy()).cuda()
    F = torch.empty(3 * nx * 16, dtype=torch.float64, device='cuda')
    strip.residual(0, U, V, F)     # pushes into the silent neighbour's region, then waits for its word
    assert strip.status()          # timed out, reported, device still alive
    assert torch.isfinite(F[:3 * nx * 8]).all()


def test_residual_peer_rejects_inconsistent_neighbours(cuda):
    from lid_driven_cavity_problem_b200 import peer_halo
    torch, native = cuda['torch'], cuda['native']
    g, X = synthetic_case(32, 32)
    buf = torch.zeros(peer_halo.region_bytes(32, 16, 1) // 8, dtype=torch.float64, device='cuda')
    grid = native.grid_from_graph(g, row_begin=8, row_count=16)
    with pytest.raises(ValueError):
        peer_halo.PeerStripResidual(grid, buf.data_ptr(), 0, buf.data_ptr(), 8, 16)
    ph = native.LdcPeerHalo()
    ph.epoch = 1
    rc = native.lib().ldc_residual_peer(grid, ph, native.ptr(buf), native.ptr(buf), native.ptr(buf),
                                        buf.data_ptr() + 8 * 4096, None, None, native.current_stream())
    assert rc != 0 and b'neighbour' in native.lib().ldc_last_error()
    rc = native.lib().ldc_residual_peer_push(grid, ph, native.ptr(buf), native.current_stream())
    assert rc != 0 and b'neighbour' in native.lib().ldc_last_error()


def test_mask_bit_exact(cuda, golden):
    from oracle import ldc_oracle as orc
    common = cuda['common']
    d = golden('jacobian_masks.npz')
    for key in sorted(k[len('indptr_'):] for k in d.files if k.startswith('indptr_')):
        nx, ny, dof = [int(v) for v in key.split('_')]
        indptr, indices = common.jacobian_mask_device(nx, ny, dof)
        assert np.array_equal(indptr.cpu().numpy(), d['indptr_' + key]), key
        assert np.array_equal(indices.cpu().numpy(), d['indices_' + key]), key
    for nx, ny in [(100, 100), (256, 256), (64, 64), (1024, 1024)]:
        indptr, indices = common.jacobian_mask_device(nx, ny, 3)
        a, b = orc.jacobian_mask_arrays(nx, ny, 3)
        assert np.array_equal(indptr.cpu().numpy(), a) and np.array_equal(indices.cpu().numpy(), b)
    m = common._calculate_jacobian_mask(500, 500, 3)   # reference smoke test
    assert m.nnz == 15732000 and m.has_sorted_indices and m.indices.dtype == np.int32
    assert np.all(m.data == 1.0)


@pytest.mark.parametrize("nx,ny,bc", [(9, 8, 100.0), (13, 9, 400.0), (4, 3, 10.0), (64, 64, 1000.0),
                                      (100, 100, 1000.0), (131, 77, 3200.0), (256, 256, 1000.0),
                                      (1024, 1024, 3200.0)])
def test_fd_jacobian_vs_oracle(cuda, nx, ny, bc):
    from oracle import ldc_oracle as orc
    torch, native = cuda['torch'], cuda['native']
    g, X = synthetic_case(nx, ny, bc=bc, seed=7 * nx + ny)
    # dummy slots carry non-zero values as they do in the middle of a Newton solve
    X[1::3][nx - 1::nx] = 0.3
    X[2::3][nx * (ny - 1):] = -0.2
    mask = orc.jacobian_mask_arrays(nx, ny, 3)
    J_ref = orc.fd_jacobian(X, g, mask=mask)
    Xd = torch.from_numpy(X).cuda()
    Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
    Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
    vals = torch.full((mask[1].size,), np.nan, dtype=torch.float64, device='cuda')
    grid = native.grid_from_graph(g)
    native.check(native.lib().ldc_fd_jacobian(grid, native.ptr(Xd), native.ptr(Ud), native.ptr(Vd),
                                              native.ptr(vals), native.current_stream()))
    J = vals.cpu().numpy()
    assert np.array_equal(J, J_ref)


@pytest.mark.parametrize("nx,ny", [(9, 8), (64, 64), (100, 100), (257, 131)])
def test_spmv_vs_oracle(cuda, nx, ny):
    from oracle import ldc_oracle as orc
    torch, native, common = cuda['torch'], cuda['native'], cuda['common']
    g, X = synthetic_case(nx, ny)
    mask = orc.jacobian_mask_arrays(nx, ny, 3)
    J = orc.fd_jacobian(X, g, mask=mask)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3 * nx * ny)
    y_ref = orc.spmv(mask, J, x)
    Jd, xd = torch.from_numpy(J).cuda(), torch.from_numpy(x).cuda()
    y = torch.empty_like(xd)
    grid = native.grid_from_graph(g)
    native.check(native.lib().ldc_spmv_mask(grid, native.ptr(Jd), native.ptr(xd), native.ptr(y),
                                            native.current_stream()))
    assert np.array_equal(y.cpu().numpy(), y_ref)
    indptr, indices = common.jacobian_mask_device(nx, ny, 3)
    y2 = torch.empty_like(xd)
    native.check(native.lib().ldc_spmv_csr(3 * nx * ny, native.ptr(indptr), native.ptr(indices), native.ptr(Jd),
                                           native.ptr(xd), native.ptr(y2), native.current_stream()))
    assert np.array_equal(y2.cpu().numpy(), y_ref)


def test_inputs_are_validated(cuda):
    crf = cuda['crf']
    g, X = synthetic_case(16, 16)
    with pytest.raises(ValueError):
        crf.residual_function(X[:-1], g)
    F32 = crf.residual_function(X.astype(np.float32), g)   # converted, not reinterpreted
    assert F32.dtype == np.float64 and F32.shape == X.shape
    X0 = X.copy()
    crf.residual_function(X, g)
    assert np.array_equal(X, X0)                            # pure


# positions of the 26 compact entries as (row dof, cell offset index k in the mask order, column dof)
_COMPACT = [(0, 0, 2), (0, 2, 1), (0, 3, 1), (0, 3, 2),
            (1, 0, 1), (1, 0, 2), (1, 1, 2), (1, 2, 1), (1, 3, 0), (1, 3, 1), (1, 3, 2), (1, 4, 0), (1, 4, 1),
            (1, 4, 2), (1, 6, 1),
            (2, 0, 2), (2, 2, 1), (2, 2, 2), (2, 3, 0), (2, 3, 1), (2, 3, 2), (2, 4, 2), (2, 5, 1), (2, 6, 0),
            (2, 6, 1), (2, 6, 2)]


@pytest.mark.parametrize("nx,ny", [(9, 8), (64, 64), (100, 100), (300, 70)])
def test_compact_jacobian_and_spmv(cuda, nx, ny):
    """the solver's 26-entry layout holds exactly the CSR numbers; every other mask entry is 0"""
    from oracle import ldc_oracle as orc
    import scipy.sparse as sp
    torch, native = cuda['torch'], cuda['native']
    g, X = synthetic_case(nx, ny, seed=3 * nx + ny)
    N = nx * ny
    indptr, indices = orc.jacobian_mask_arrays(nx, ny, 3)
    J_ref = orc.fd_jacobian(X, g, mask=(indptr, indices))
    A = sp.csr_matrix((J_ref, indices, indptr), shape=(3 * N, 3 * N))
    Xd = torch.from_numpy(X).cuda()
    Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
    Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
    Jc = torch.full((26 * N,), np.nan, dtype=torch.float64, device='cuda')
    grid = native.grid_from_graph(g)
    native.check(native.lib().ldc_fd_jacobian_compact(grid, native.ptr(Xd), native.ptr(Ud), native.ptr(Vd),
                                                      native.ptr(Jc), native.current_stream()))
    Jc = Jc.cpu().numpy().reshape(26, N)
    offs = [-nx, -nx + 1, -1, 0, 1, nx - 1, nx]
    rebuilt = sp.lil_matrix((3 * N, 3 * N))
    cells = np.arange(N)
    for k, (d, o, e) in enumerate(_COMPACT):
        cc = cells + offs[o]
        ok = (cc >= 0) & (cc < N)
        assert np.all(Jc[k][~ok] == 0.0)
        rows, cols = 3 * cells[ok] + d, 3 * cc[ok] + e
        assert np.array_equal(Jc[k][ok], np.asarray(A[rows, cols]).ravel()), k
        rebuilt[rows, cols] = Jc[k][ok]
    assert (abs(rebuilt.tocsr() - A)).max() == 0.0          # nothing outside the 26 entries
    x = np.random.default_rng(1).standard_normal(3 * N)
    xd = torch.from_numpy(x).cuda()
    y = torch.empty_like(xd)
    native.check(native.lib().ldc_spmv_compact(grid, native.ptr(torch.from_numpy(Jc.ravel()).cuda()), native.ptr(xd),
                                               native.ptr(y), native.current_stream()))
    y_ref = A @ x
    assert np.abs(y.cpu().numpy() - y_ref).max() <= 1e-13 * np.abs(y_ref).max()


def test_compact_jacobian_and_spmv_row_strips(cuda):
    """row-strip decomposition of the Jacobian assembly and the SpMV: every strip (with its halo
    rows of X / x) reproduces its slice of the full-grid result bit for bit"""
    torch, native = cuda['torch'], cuda['native']
    lib = native.lib()
    nx, ny = 64, 50
    g, X = synthetic_case(nx, ny)
    N = nx * ny
    st = native.current_stream()
    Xd = torch.from_numpy(X).cuda()
    Ud = torch.from_numpy(g.ns_x_mesh.phi_old).cuda()
    Vd = torch.from_numpy(g.ns_y_mesh.phi_old).cuda()
    Jfull = torch.empty(26 * N, dtype=torch.float64, device='cuda')
    grid = native.grid_from_graph(g)
    native.check(lib.ldc_fd_jacobian_compact(grid, native.ptr(Xd), native.ptr(Ud), native.ptr(Vd), native.ptr(Jfull), st))
    x = np.random.default_rng(2).standard_normal(3 * N)
    xd = torch.from_numpy(x).cuda()
    yfull = torch.empty_like(xd)
    native.check(lib.ldc_spmv_compact(grid, native.ptr(Jfull), native.ptr(xd), native.ptr(yfull), st))
    Jfull = Jfull.cpu().numpy().reshape(26, ny, nx)
    yfull = yfull.cpu().numpy()
    U_old = g.ns_x_mesh.phi_old.reshape(ny, nx - 1)
    V_old = g.ns_y_mesh.phi_old.reshape(ny - 1, nx)
    bounds = [0, 13, 14, 32, 50]
    for j0, j1 in zip(bounds[:-1], bounds[1:]):
        lo, hi = max(j0 - 1, 0), min(j1 + 1, ny)
        rows = j1 - j0
        Xl = torch.from_numpy(X[3 * nx * lo:3 * nx * hi].copy()).cuda()
        xl = torch.from_numpy(x[3 * nx * lo:3 * nx * hi].copy()).cuda()
        Ul = torch.from_numpy(U_old[j0:j1].copy().ravel()).cuda()
        Vl = torch.from_numpy(np.r_[V_old[j0:min(j1, ny - 1)].ravel(), np.zeros(nx)]).cuda()
        sg = native.grid_from_graph(g, row_begin=j0, row_count=rows)
        off = 8 * 3 * nx * (j0 - lo)
        Jl = torch.full((26 * nx * rows,), np.nan, dtype=torch.float64, device='cuda')
        native.check(lib.ldc_fd_jacobian_compact(sg, Xl.data_ptr() + off, native.ptr(Ul), native.ptr(Vl), native.ptr(Jl), st))
        assert np.array_equal(Jl.cpu().numpy().reshape(26, rows, nx), Jfull[:, j0:j1, :]), (j0, j1)
        yl = torch.empty(3 * nx * rows, dtype=torch.float64, device='cuda')
        native.check(lib.ldc_spmv_compact(sg, native.ptr(Jl), xl.data_ptr() + off, native.ptr(yl), st))
        assert np.array_equal(yl.cpu().numpy(), yfull[3 * nx * j0:3 * nx * j1]), (j0, j1)


def test_ensemble_512x64_residual_full_size(cuda):
    """BASELINE config 4 at full size: 512 cavities of 64x64, Re log-spaced 100..10000, one launch"""
    from oracle import ldc_oracle as orc
    torch, native = cuda['torch'], cuda['native']
    B, n = 512, 64
    bcs = 100.0 * 100.0 ** (np.arange(B) / (B - 1))
    Xs, Us, Vs, Fs = [], [], [], []
    for b in range(B):
        g, X = synthetic_case(n, n, bc=float(bcs[b]), seed=b)
        Xs.append(X); Us.append(g.ns_x_mesh.phi_old); Vs.append(g.ns_y_mesh.phi_old)
        Fs.append(orc.residual_function(X, g))
    g, _ = synthetic_case(n, n)
    grid = native.grid_from_graph(g, batch=B)
    bc_dev = torch.from_numpy(bcs).cuda()
    grid.bc_dev = bc_dev.data_ptr()
    Xd = torch.from_numpy(np.concatenate(Xs)).cuda()
    Ud = torch.from_numpy(np.concatenate(Us)).cuda()
    Vd = torch.from_numpy(np.concatenate(Vs)).cuda()
    F = torch.empty_like(Xd)
    n2 = torch.zeros(B, dtype=torch.float64, device='cuda')
    work = torch.empty(int(native.lib().ldc_residual_work_doubles(grid)), dtype=torch.float64, device='cuda')
    native.check(native.lib().ldc_residual(grid, native.ptr(Xd), native.ptr(Ud), native.ptr(Vd), native.ptr(F),
                                           native.ptr(n2), native.ptr(work), 0, native.current_stream()))
    Fh = F.cpu().numpy()
    assert np.array_equal(Fh, np.concatenate(Fs))
    ref_n2 = np.array([np.dot(f, f) for f in Fs])
    assert np.abs(n2.cpu().numpy() - ref_n2).max() <= 1e-12 * ref_n2.max()
