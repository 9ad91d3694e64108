"""CPU prototype (numpy/scipy) of the preconditioner the CUDA solver implements:
additive cell-block ("Vanka") relaxation + geometric multigrid with rediscretised coarse
Jacobians.  Development tool only -- uses the oracle for residual/Jacobian."""
import sys, time, os
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from oracle import ldc_oracle as orc
from conftest import make_graph


class Level:
    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny
        self.N = nx * ny
        self.n = 3 * self.N


def idx(nx, i, j, d):
    return 3 * (i + nx * j) + d


def build_vanka(L, J):
    """extract per-cell coefficients; returns dict of arrays shaped (ny,nx)"""
    nx, ny = L.nx, L.ny
    I, Jj = np.meshgrid(np.arange(nx), np.arange(ny))
    c = I + nx * Jj
    Jc = J.tocsr()
    def get(rows, cols, mask):
        out = np.zeros(rows.shape)
        r, cc = rows[mask], cols[mask]
        out[mask] = np.asarray(Jc[r, cc]).ravel()
        return out
    W, E, S, Nn = I > 0, I < nx - 1, Jj > 0, Jj < ny - 1
    pc = 3 * c
    K = {}
    # east face unknown of this cell: U slot (c,1); west: (c-1,1); north V: (c,2); south: (c-nx,2)
    faces = {'e': (3 * c + 1, E), 'w': (3 * (c - 1) + 1, W), 'n': (3 * c + 2, Nn), 's': (3 * (c - nx) + 2, S)}
    for k, (slot, m) in faces.items():
        K['a' + k] = get(slot, slot, m)          # momentum diagonal
        K['g' + k] = get(slot, pc, m)            # d mom_k / d p_c
        K['d' + k] = get(pc, slot, m)            # d cont_c / d vel_k
        K['m' + k] = m
        K['slot' + k] = slot
    K['c'] = c
    # dummy slots
    dm = np.zeros(L.n, bool)
    dm[(3 * c + 1)[~E]] = True
    dm[(3 * c + 2)[~Nn]] = True
    K['dummy'] = dm
    K['diag'] = Jc.diagonal()
    return K


def vanka_sweep(L, J, K, x, b, om_u, om_p):
    r = b - J @ x
    nx, ny = L.nx, L.ny
    c = K['c']
    num = -r[3 * c]
    den = np.zeros_like(num)
    for k in 'ewns':
        m = K['m' + k]
        a = np.where(m, K['a' + k], 1.0)
        rk = np.where(m, r[np.where(m, K['slot' + k], 0)], 0.0)
        num = num + np.where(m, K['d' + k] * rk / a, 0.0)
        den = den + np.where(m, K['d' + k] * K['g' + k] / a, 0.0)
    dp = num / den
    dx = np.zeros_like(x)
    dx[3 * c.ravel()] = om_p * dp.ravel()
    for k in 'ewns':
        m = K['m' + k]
        a = np.where(m, K['a' + k], 1.0)
        slot = K['slot' + k]
        rk = np.where(m, r[np.where(m, slot, 0)], 0.0)
        du = (rk - K['g' + k] * dp) / a
        np.add.at(dx, slot[m], 0.5 * om_u * du[m])
    dm = K['dummy']
    dx[dm] = r[dm] / K['diag'][dm]
    return x + dx


def restrict_state(Lf, Lc, X):
    nx, ny = Lf.nx, Lf.ny
    U = X[1::3].reshape(ny, nx); V = X[2::3].reshape(ny, nx)
    Uc = 0.5 * (U[0::2, 1::2] + U[1::2, 1::2])
    Vc = 0.5 * (V[1::2, 0::2] + V[1::2, 1::2])
    Xc = np.zeros(Lc.n)
    Uc = Uc.copy(); Uc[:, -1] = 0.0
    Vc = Vc.copy(); Vc[-1, :] = 0.0
    Pf = X[0::3].reshape(ny, nx)
    Xc[0::3] = (0.25 * (Pf[0::2, 0::2] + Pf[1::2, 0::2] + Pf[0::2, 1::2] + Pf[1::2, 1::2])).ravel()
    Xc[1::3] = Uc.ravel(); Xc[2::3] = Vc.ravel()
    return Xc


def build_transfer(Lf, Lc):
    """R (coarse n x fine n), residual restriction; P = R^T variants"""
    nxf, nyf, nxc, nyc = Lf.nx, Lf.ny, Lc.nx, Lc.ny
    rows, cols, vals = [], [], []
    for Jc in range(nyc):
        for Ic in range(nxc):
            # continuity / pressure: 4 children
            for dj in (0, 1):
                for di in (0, 1):
                    rows.append(idx(nxc, Ic, Jc, 0)); cols.append(idx(nxf, 2 * Ic + di, 2 * Jc + dj, 0)); vals.append(1.0)
            # U
            if Ic < nxc - 1:
                for dj in (0, 1):
                    for di, w in ((0, .5), (1, 1.), (2, .5)):
                        rows.append(idx(nxc, Ic, Jc, 1)); cols.append(idx(nxf, 2 * Ic + di, 2 * Jc + dj, 1)); vals.append(w)
            if Jc < nyc - 1:
                for di in (0, 1):
                    for dj, w in ((0, .5), (1, 1.), (2, .5)):
                        rows.append(idx(nxc, Ic, Jc, 2)); cols.append(idx(nxf, 2 * Ic + di, 2 * Jc + dj, 2)); vals.append(w)
    R = sp.csr_matrix((vals, (rows, cols)), shape=(Lc.n, Lf.n))
    # prolongation: velocities R^T (x-linear, y-constant); pressure: injection of parent = R^T on P block
    P = R.T.tocsr()
    return R, P


class MG:
    def __init__(self, graph, X, min_size=4, nu=(2, 2), om_u=0.7, om_p=0.7, coarse='direct', max_levels=20):
        self.levels = []
        nx, ny = graph.pressure_mesh.nx, graph.pressure_mesh.ny
        g = graph
        Xl = X
        while True:
            L = Level(nx, ny)
            mask = orc.jacobian_mask_arrays(nx, ny, 3)
            vals = orc.fd_jacobian(Xl, g, mask=mask)
            L.J = sp.csr_matrix((vals, mask[1], mask[0]), shape=(L.n, L.n))
            L.K = build_vanka(L, L.J)
            L.g = g
            self.levels.append(L)
            if nx % 2 or ny % 2 or nx // 2 < min_size or ny // 2 < min_size or len(self.levels) >= max_levels:
                break
            Lc = Level(nx // 2, ny // 2)
            L.R, L.P = build_transfer(L, Lc)
            Xl = restrict_state(L, Lc, Xl)
            gc = make_graph([1.0, 1.0, nx // 2, ny // 2, g.dt, g.rho, g.mi, g.bc])
            gc.dx, gc.dy = 2 * g.dx, 2 * g.dy
            gc.ns_x_mesh.phi_old = np.zeros(len(gc.ns_x_mesh)); gc.ns_y_mesh.phi_old = np.zeros(len(gc.ns_y_mesh))
            g = gc
            nx, ny = nx // 2, ny // 2
        self.nu, self.om_u, self.om_p, self.coarse = nu, om_u, om_p, coarse
        Lc = self.levels[-1]
        if coarse == 'direct':
            A = Lc.J.tolil(); A[0, :] = 0; A[0, 0] = 1.0
            Lc.lu = spla.splu(A.tocsc())

    def vcycle(self, l, b):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            if self.coarse == 'direct':
                bb = b.copy(); bb[0] = 0
                return L.lu.solve(bb)
            x = np.zeros(L.n)
            for _ in range(self.coarse):
                x = vanka_sweep(L, L.J, L.K, x, b, self.om_u, self.om_p)
            return x
        x = np.zeros(L.n)
        for _ in range(self.nu[0]):
            x = vanka_sweep(L, L.J, L.K, x, b, self.om_u, self.om_p)
        r = b - L.J @ x
        xc = self.vcycle(l + 1, L.R @ r)
        x = x + L.P @ xc
        for _ in range(self.nu[1]):
            x = vanka_sweep(L, L.J, L.K, x, b, self.om_u, self.om_p)
        return x


def fgmres(A, b, M, restart=30, rtol=1e-5, max_it=500):
    """right-preconditioned GMRES(restart) (flexible form); returns x, its, relres"""
    n = b.size
    x = np.zeros(n)
    bnorm = np.linalg.norm(b)
    its = 0
    r = b.copy()
    beta = np.linalg.norm(r)
    hist = [beta]
    while its < max_it:
        V = np.zeros((restart + 1, n)); Z = np.zeros((restart, n))
        H = np.zeros((restart + 1, restart))
        V[0] = r / beta
        g = np.zeros(restart + 1); g[0] = beta
        cs = np.zeros(restart); sn = np.zeros(restart)
        k = 0
        done = False
        for k in range(restart):
            Z[k] = M(V[k])
            w = A @ Z[k]
            h = V[:k + 1] @ w
            w = w - h @ V[:k + 1]
            H[:k + 1, k] = h
            H[k + 1, k] = np.linalg.norm(w)
            V[k + 1] = w / H[k + 1, k]
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]; H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k]); cs[k] = H[k, k] / d; sn[k] = H[k + 1, k] / d
            H[k, k] = d; H[k + 1, k] = 0
            g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k]
            its += 1
            hist.append(abs(g[k + 1]))
            if abs(g[k + 1]) <= rtol * bnorm or its >= max_it:
                done = True
                break
        kk = k + 1
        y = np.linalg.solve(np.triu(H[:kk, :kk]), g[:kk])
        x = x + y @ Z[:kk]
        r = b - A @ x
        beta = np.linalg.norm(r)
        if done and beta <= rtol * bnorm * 1.001 or its >= max_it:
            break
    return x, its, beta / bnorm, hist


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=64)
    ap.add_argument('--re', type=float, default=1000.0)
    ap.add_argument('--mode', default='mg')
    ap.add_argument('--omu', type=float, default=0.7)
    ap.add_argument('--omp', type=float, default=0.7)
    ap.add_argument('--nu', type=int, default=2)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--coarse', default='direct')
    ap.add_argument('--min', type=int, default=4)
    args = ap.parse_args()
    n = args.n
    g = make_graph([1.0, 1.0, n, n, 1e-2, 1.0, 1.0, args.re])
    # Newton with the prototype linear solver, reference tolerances
    t = 0.0
    for step in range(args.steps):
        g.ns_x_mesh.phi_old = g.ns_x_mesh.phi.copy(); g.ns_y_mesh.phi_old = g.ns_y_mesh.phi.copy()
        X = orc.create_X(g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi, g)
        F = orc.residual_function(X, g, nthreads=0)
        f0 = np.linalg.norm(F)
        for it in range(50):
            t0 = time.time()
            coarse = args.coarse if args.coarse == 'direct' else int(args.coarse)
            mg = MG(g, X, nu=(args.nu, args.nu), om_u=args.omu, om_p=args.omp, coarse=coarse, min_size=args.min,
                    max_levels=1 if args.mode == 'vanka' else 20)
            A = mg.levels[0].J
            if args.mode == 'vanka':
                L0 = mg.levels[0]
                def M(v):
                    x = np.zeros_like(v)
                    for _ in range(args.nu):
                        x = vanka_sweep(L0, L0.J, L0.K, x, v, args.omu, args.omp)
                    return x
            else:
                M = lambda v: mg.vcycle(0, v)
            y, its, rel, hist = fgmres(A, F, M, max_it=3000 if args.mode == 'vanka' else 300)
            X = X - y
            F = orc.residual_function(X, g, nthreads=0)
            fn = np.linalg.norm(F)
            print('step %d newton %d: levels=%d gmres its=%d rel=%.2e  |F|=%.3e  (%.1fs)' % (
                step, it, len(mg.levels), its, rel, fn, time.time() - t0), flush=True)
            if fn < 1e-4 or fn <= 1e-4 * f0 or np.linalg.norm(y) < 1e-4 * np.linalg.norm(X):
                break
        U, V, P = orc.recover_X(X, g)
        g.ns_x_mesh.phi, g.ns_y_mesh.phi, g.pressure_mesh.phi = U, V, P
        g.dt *= 1.2
