// Device-resident Newton-Krylov solve of one pseudo-time step (sm_100a).
//
// Replaces PetscSolverWrapper.solve / _setup_snes / _setup_options
// (nonlinear_solver/petsc_solver_wrapper.py:35-145), i.e. PETSc's SNES newtonls with line
// search 'basic' (full steps), Jacobian by coloured finite differences, KSP GMRES(30):
//   * state, residual, Jacobian values, Krylov basis and the multigrid hierarchy live in HBM
//     for the whole run; per Newton iteration the host reads back three norms per cavity,
//     per Krylov iteration one residual estimate per cavity (a few bytes);
//   * mask / colouring / workspace are built once per solver (the reference rebuilds its SNES,
//     mask and colouring every time step because `_first_run` is never cleared,
//     petsc_solver_wrapper.py:39,71-72);
//   * Newton tolerances, reason codes and the order of the convergence tests follow
//     SNESConvergedDefault (rtol=atol=stol=1e-4, max_it=50, petsc_solver_wrapper.py:135);
//   * the linear solve is restarted flexible GMRES, right-preconditioned by the multigrid
//     V-cycle of precond.cu, converged on the TRUE residual ||F - J y|| <= max(rtol*||F||, atol)
//     (PETSc's default is left preconditioning on the preconditioned residual; only the
//     converged state has to agree with the reference).
// A batch of independent cavities (Reynolds ensemble) advances in lock step; finished cavities
// are masked out on the device.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "solver_internal.cuh"

namespace ldc {

// per-instance GMRES scalars on the device
struct KrylovSmall {
    double *H;      // [batch][(m+1)*m] column major
    double *cs, *sn;  // [batch][m]
    double *gv;     // [batch][m+1]
    double *yv;     // [batch][m]
    double *hcol;   // [batch][k+1] packed: dots of the current column against v_0..v_k
    double *nrm2;   // [batch]        scratch for reduced squared norms
    double *beta;   // [batch]        residual norm at cycle start
    double *res;    // [batch]        current residual estimate
    double *hnext;  // [batch]        h_{k+1,k} before rotation
    double *hcol2;  // [batch][k+1] packed: second (re-orthogonalisation) pass
};

// h += h2 (packed columns of all instances)
__global__ void add_columns_kernel(double *h, const double *h2, int total)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) h[t] += h2[t];
}

__global__ void krylov_begin_cycle_kernel(KrylovSmall ks, int m, int batch, const int32_t *run)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (run && !run[b]) return;
    double *gv = ks.gv + (long)b * (m + 1);
    for (int i = 0; i <= m; ++i) gv[i] = 0.0;
    gv[0] = ks.beta[b];
    ks.res[b] = ks.beta[b];
}

// sqrt of a reduced squared norm (per instance)
__global__ void sqrt_kernel(const double *in, double *out, int batch, const int32_t *run)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (run && !run[b]) return;
    out[b] = sqrt(in[b]);
}

// column k: store h, apply the stored Givens rotations, create the new one, update g
__global__ void krylov_givens_kernel(KrylovSmall ks, int m, int k, int batch, const int32_t *run)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (run && !run[b]) return;
    double *H = ks.H + (long)b * (m + 1) * m + (long)k * (m + 1);
    double *cs = ks.cs + (long)b * m, *sn = ks.sn + (long)b * m;
    double *gv = ks.gv + (long)b * (m + 1);
    const double *h = ks.hcol + (long)b * (k + 1);  // packed [batch][k+1] by vec_reduce_partials
    const double hk1 = sqrt(ks.nrm2[b]);
    ks.hnext[b] = hk1;
    for (int i = 0; i <= k; ++i) H[i] = h[i];
    H[k + 1] = hk1;
    for (int i = 0; i < k; ++i) {
        const double a = H[i], c = H[i + 1];
        H[i] = cs[i] * a + sn[i] * c;
        H[i + 1] = -sn[i] * a + cs[i] * c;
    }
    const double a = H[k], c = H[k + 1];
    const double d = sqrt(a * a + c * c);
    if (d == 0.0) { cs[k] = 1.0; sn[k] = 0.0; }
    else { cs[k] = a / d; sn[k] = c / d; }
    H[k] = cs[k] * a + sn[k] * c;
    H[k + 1] = 0.0;
    gv[k + 1] = -sn[k] * gv[k];
    gv[k] = cs[k] * gv[k];
    ks.res[b] = fabs(gv[k + 1]);
}

// back substitution H(0:kk,0:kk) y = g(0:kk)
__global__ void krylov_solve_y_kernel(KrylovSmall ks, int m, int kk, int batch, const int32_t *run)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (run && !run[b]) return;
    const double *H = ks.H + (long)b * (m + 1) * m;
    const double *gv = ks.gv + (long)b * (m + 1);
    double *y = ks.yv + (long)b * m;
    for (int i = kk - 1; i >= 0; --i) {
        double s = gv[i];
        for (int j = i + 1; j < kk; ++j) s -= H[(long)j * (m + 1) + i] * y[j];
        const double d = H[(long)i * (m + 1) + i];
        y[i] = (d != 0.0) ? s / d : 0.0;
    }
}

// U_old, V_old <- the velocity slots of X (masked per instance)
__global__ void __launch_bounds__(256)
extract_old_kernel(int nx, int ny, int v_rows, const double *__restrict__ X, double *__restrict__ U_old,
                   double *__restrict__ V_old, const int32_t *__restrict__ run)
{
    // ny = rows held by this solver (grid rows, or the rows of a strip); v_rows = how many of them
    // carry a real V (all but the top grid row)
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const long N = (long)nx * ny;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    const int j = (int)(c / nx), i = (int)(c - (long)j * nx);
    const double *x = X + 3 * ((long)b * N + c);
    if (i < nx - 1) U_old[(long)b * (nx - 1) * ny + i + (long)(nx - 1) * j] = x[1];
    if (j < v_rows) V_old[(long)b * nx * v_rows + c] = x[2];
}

}  // namespace ldc

using namespace ldc;

struct ldc_solver {
    ldc_grid g;
    ldc_solver_options opt;
    int batch, m;
    long N, n, nnz;
    std::vector<void *> dev_allocs, host_allocs;
    // state
    double *X, *X_prev, *F, *Y, *U_old, *V_old;
    double *dt_dev, *bc_dev;
    std::vector<MgLevel> levels;
    // Krylov
    double *V, *Z, *w, *rres;
    double *partials, *res_work;
    long partial_doubles;
    KrylovSmall ks;
    int32_t *run_dev, *fin_dev;
    double *scal_dev;   // [3*batch] fnorm2, ynorm2, xnorm2
    // pinned host mirrors
    double *scal_host, *res_host, *du_host;
    int32_t *mask_host;
    std::vector<double> dt_host, bc_host;
    long total_linear_its;
    long lockstep_krylov_its, lockstep_newton_its, kernel_launch_estimate;
    // private stream (graph capture is illegal on the legacy default stream) and one CUDA graph per
    // position k in the restart cycle: every kernel argument of Krylov iteration k is fixed for the
    // life of the solver (basis vectors V_k/Z_k, level buffers, device scalars), so the ~200 small
    // launches of one preconditioned iteration replay as a single graph launch.
    cudaStream_t stream;
    cudaEvent_t ev_in;
    // row-strip decomposition
    bool strip, ghost_lo, ghost_hi, has_comm;
    bool keep_diverged;   // ldc_solver_set_keep_diverged: no roll-back of diverged instances
    bool mixed;   // multigrid cycle in fp32
    ldc_comm_ops comm;
    int v_rows;
    double *X_alloc, *xpad_alloc, *xpad;
    long zs;   // distance between preconditioned basis vectors Z_k (each padded by a halo row on both sides)
    bool use_graphs;
    std::vector<cudaGraphExec_t> iter_graph;

    template <typename T>
    int dalloc(T **p, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
        if (e != cudaSuccess) {
            set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return LDC_ERR_ALLOC;
        }
        dev_allocs.push_back(q);
        *p = (T *)q;
        return LDC_OK;
    }
    template <typename T>
    int halloc(T **p, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = cudaMallocHost(&q, count * sizeof(T) + 64);
        if (e != cudaSuccess) {
            set_error("cudaMallocHost failed: %s", cudaGetErrorString(e));
            return LDC_ERR_ALLOC;
        }
        host_allocs.push_back(q);
        *p = (T *)q;
        return LDC_OK;
    }
};

#define TRY(expr)                    \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != LDC_OK) return rc__; \
    } while (0)

extern "C" void ldc_solver_default_options(ldc_solver_options *o)
{
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->snes_rtol = o->snes_atol = o->snes_stol = 1e-4;  // petsc_solver_wrapper.py:135
    o->snes_max_it = 50;
    o->snes_max_funcs = 10000;
    o->snes_divtol = 1e4;   // PETSc -snes_divergence_tolerance default
    o->ksp_rtol = 1e-3;     // inexact Newton: the nonlinear tolerance is 1e-4, solving every linear system to
                            // PETSc's default 1e-5 costs 25-35% more Krylov work for the same Newton path
    o->ksp_atol = 1e-50;
    o->ksp_dtol = 1e5;
    o->ksp_stagnation = 0.9;
    o->ksp_max_it = 150;    // PETSc: 10000 (ILU).  A multigrid-preconditioned solve that needs more is a
                            // Newton step about to diverge: failing fast (-> dt halving) is 1.4-2x cheaper
    o->gmres_restart = 30;
    o->gmres_refine = 0;    // PETSc's default (classical Gram-Schmidt, never refine)
    o->use_cuda_graphs = 1;
    o->pc_fp32 = 1;
    o->pc_type = LDC_PC_MG_VANKA;
    o->pc_sweeps = 2;
    o->mg_levels = 0;
    o->mg_coarse_sweeps = 10;
    o->mg_min_size = 4;
    o->pc_omega_u = 0.8;    // measured on 1024^2 Re=1000: 0.8 is ~20% fewer iterations than 0.7; 1.0 diverges
    o->pc_omega_p = 0.8;
    o->residual_variant = 0;
}

static ldc_grid level_grid(const ldc_solver *s, const MgLevel &l)
{
    ldc_grid g = s->g;
    g.nx = l.nx;
    g.ny = l.ny_global;
    g.row_begin = l.row_begin;
    g.row_count = l.ny;
    g.dx = l.dx;
    g.dy = l.dy;
    g.dt_dev = s->dt_dev;
    g.bc_dev = s->bc_dev;
    return g;
}

namespace ldc {
// Shape of the multigrid hierarchy of the cavity / row strip `g`: number of levels and the first
// replicated level (-1: none).  Host arithmetic only (ldc_solver_plan_levels exposes it so that
// the partition rules can be tested without a device).
static void plan_levels(const ldc_grid *g, const ldc_solver_options &o, int *nlev_out, int *l_rep_out)
{
    long kRepCells = 32768;
    if (const char *env = getenv("LDC_REP_CELLS")) kRepCells = atol(env) > 0 ? atol(env) : kRepCells;
    const bool strip = !(g->row_begin == 0 && g->row_count == g->ny);
    int nlev = 1, l_rep = -1;
    if (o.pc_type == LDC_PC_MG_VANKA) {
        int nx = g->nx, ny = g->row_count, rb = g->row_begin, nyg = g->ny;
        const int min_size = o.mg_min_size > 2 ? o.mg_min_size : 3;
        bool rep = false;
        while ((o.mg_levels <= 0 || nlev < o.mg_levels) && nx % 2 == 0 && nyg % 2 == 0 && nx / 2 >= min_size) {
            if (!rep) {
                if (ny % 2 != 0 || rb % 2 != 0) break;
                if (strip && ((long)(nx / 2) * (nyg / 2) <= kRepCells || ny / 2 < 2)) {
                    if (ny / 2 < 1) break;
                    rep = true;
                    l_rep = nlev;
                } else if (!strip && ny / 2 < min_size) break;
            } else if (nyg / 2 < min_size) break;
            nx /= 2; ny /= 2; rb /= 2; nyg /= 2; ++nlev;
        }
    }
    *nlev_out = nlev;
    *l_rep_out = l_rep;
}
}  // namespace ldc

extern "C" int ldc_solver_plan_levels(const ldc_grid *g, const ldc_solver_options *opt, int32_t *levels_out,
                                      int32_t *replicated_from_out)
{
    if (int rc = validate_grid(g, false)) return rc;
    LDC_REQUIRE(opt && levels_out && replicated_from_out, "ldc_solver_plan_levels: null argument");
    int nlev = 1, l_rep = -1;
    ldc::plan_levels(g, *opt, &nlev, &l_rep);
    *levels_out = nlev;
    *replicated_from_out = l_rep;
    return LDC_OK;
}

extern "C" int ldc_solver_create(const ldc_grid *g, const ldc_solver_options *opt, ldc_solver **out)
{
    LDC_REQUIRE(out != nullptr, "ldc_solver_create: out is NULL");
    *out = nullptr;
    TRY(validate_grid(g, false));
    ldc_solver_options o;
    if (opt) o = *opt; else ldc_solver_default_options(&o);
    LDC_REQUIRE(o.gmres_restart >= 1 && o.gmres_restart <= kMaxRestart, "gmres_restart must be in [1,%d]", kMaxRestart);
    LDC_REQUIRE(o.pc_type >= LDC_PC_NONE && o.pc_type <= LDC_PC_MG_VANKA, "unknown pc_type %d", o.pc_type);
    LDC_REQUIRE(o.pc_sweeps >= 1 && o.mg_coarse_sweeps >= 1, "sweep counts must be >= 1");
    ldc_solver *s = new ldc_solver();
    s->g = *g;
    s->opt = o;
    s->batch = g->batch;
    s->m = o.gmres_restart;
    s->strip = !(g->row_begin == 0 && g->row_count == g->ny);
    s->ghost_lo = g->row_begin > 0;
    s->ghost_hi = g->row_begin + g->row_count < g->ny;
    s->has_comm = false;
    s->keep_diverged = false;
    s->mixed = (o.pc_fp32 != 0) && o.pc_type != LDC_PC_NONE;   // strips: revisited once the levels are planned
    memset(&s->comm, 0, sizeof(s->comm));
    s->v_rows = (g->row_begin + g->row_count < g->ny ? g->row_count : g->ny - 1 - g->row_begin);
    if (s->v_rows < 0) s->v_rows = 0;
    s->N = (long)g->nx * g->row_count;   // cells held by this solver (a whole cavity or its row strip)
    s->n = 3 * s->N;
    s->nnz = ldc_mask_nnz(g->nx, g->ny, 3);
    s->total_linear_its = 0;
    s->lockstep_krylov_its = s->lockstep_newton_its = s->kernel_launch_estimate = 0;
    s->stream = nullptr;
    s->ev_in = nullptr;
    s->use_graphs = (o.use_cuda_graphs != 0) && (getenv("LDC_NO_GRAPHS") == nullptr);
    s->iter_graph.assign(s->m, nullptr);
    const int B = s->batch, m = s->m;
    const long n = s->n;
    int rc = LDC_OK;
#define A(expr) do { if (rc == LDC_OK) rc = (expr); } while (0)
    // the state and the SpMV input of a strip carry one ghost row below and above
    const long row = 3L * g->nx;
    A(s->dalloc(&s->X_alloc, (size_t)B * n + 2 * row));
    s->X = s->X_alloc ? s->X_alloc + row : nullptr;
    A(s->dalloc(&s->xpad_alloc, (size_t)n + 2 * row));
    s->xpad = s->xpad_alloc ? s->xpad_alloc + row : nullptr;
    A(s->dalloc(&s->X_prev, (size_t)B * n));
    A(s->dalloc(&s->F, (size_t)B * n));
    A(s->dalloc(&s->Y, (size_t)B * n));
    // U_old / V_old also stand in as the "old" fields of every coarse level's FD Jacobian (the old
    // values drop out of the differences): size them for the largest level.  A REPLICATED coarse
    // level holds the whole coarse grid and can be larger than a thin strip's own fine rows (256^2
    // over 8 ranks: 8192 fine cells per strip, 16384 cells on the replicated 128^2 level).
    {
        size_t old_cells = (size_t)g->nx * g->row_count;
        if (!(g->row_begin == 0 && g->row_count == g->ny)) {
            const size_t coarse_whole = ((size_t)g->nx / 2) * ((size_t)g->ny / 2);
            if (coarse_whole > old_cells) old_cells = coarse_whole;
        }
        A(s->dalloc(&s->U_old, (size_t)B * old_cells));
        A(s->dalloc(&s->V_old, (size_t)B * old_cells));
        if (rc == LDC_OK) {
            cudaMemset(s->U_old, 0, sizeof(double) * B * old_cells);
            cudaMemset(s->V_old, 0, sizeof(double) * B * old_cells);
        }
    }
    A(s->dalloc(&s->dt_dev, B));
    A(s->dalloc(&s->bc_dev, B));
    // multigrid hierarchy.  A strip coarsens its own rows; strip boundaries stay aligned because the
    // partition is a multiple of 2^levels rows (strip_partition).
    // A strip hierarchy is distributed on its fine levels and REPLICATED below: once a level is
    // small (<= kRepCells cells in the whole grid) or the strip cannot be halved any more, its
    // right-hand side is all-gathered and every rank continues the cycle on the full coarse grid.
    // That keeps the latency-bound coarse levels free of halo exchanges and makes them identical to
    // the single-GPU hierarchy.
    int nlev = 1, l_rep = -1;
    plan_levels(g, o, &nlev, &l_rep);
    if (s->strip && s->mixed) {
        // fp32 vectors of a distributed level travel through the double-typed transport: a halo
        // row of 3*nx floats as 3*nx/2 doubles, a slice of a replicated level likewise.  Needs even
        // counts (true for every grid that can be coarsened at all); otherwise the cycle stays fp64.
        for (int l = 0; l < nlev; ++l) {
            const long nxl = g->nx >> l;
            const bool rep = l_rep >= 0 && l >= l_rep;
            if ((!rep || l == l_rep) && nxl % 2 != 0) s->mixed = false;   // also keeps the rows 8-byte aligned
        }
    }
    s->levels.resize(nlev);
    for (int l = 0; l < nlev && rc == LDC_OK; ++l) {
        MgLevel &L = s->levels[l];
        L.replicated = (l_rep >= 0 && l >= l_rep);
        L.nx = g->nx >> l;
        L.ny_global = g->ny >> l;
        L.slice_begin = g->row_begin >> l;
        L.slice_rows = g->row_count >> l;
        if (L.replicated) {
            L.ny = L.ny_global;
            L.row_begin = 0;
            L.ghost_lo = L.ghost_hi = false;
        } else {
            L.ny = g->row_count >> l;
            L.row_begin = g->row_begin >> l;
            L.ghost_lo = s->ghost_lo;
            L.ghost_hi = s->ghost_hi;
        }
        L.v_rows = (L.row_begin + L.ny < L.ny_global) ? L.ny : L.ny_global - 1 - L.row_begin;
        if (L.v_rows < 0) L.v_rows = 0;
        L.dx = g->dx * (double)(1 << l);
        L.dy = g->dy * (double)(1 << l);
        L.N = (long)L.nx * L.ny;
        L.n = 3 * L.N;
        L.nnz = 0;
        L.X = nullptr; L.x = L.b = nullptr;
        const long lrow = 3L * L.nx;   // every level vector has a spare (halo) row on both sides
        auto padded = [&](double **p) {
            double *q = nullptr;
            int r2 = s->dalloc(&q, (size_t)B * L.n + 2 * lrow);
            if (r2 == LDC_OK) {
                cudaMemset(q, 0, sizeof(double) * ((size_t)B * L.n + 2 * lrow));
                *p = q + lrow;
            }
            return r2;
        };
        if (l == 0) L.X = s->X;
        else {
            A(padded(&L.X));
            A(padded(&L.x));
            A(padded(&L.b));
        }
        A(s->dalloc(&L.J, (size_t)B * 26 * L.N));
        A(s->dalloc(&L.diag, (size_t)B * L.n));
        A(padded(&L.r));
        A(s->dalloc(&L.dp, (size_t)B * L.N));
        L.Jf = L.diagf = L.xf = L.bf = L.rf = L.dpf = nullptr;
        if (s->mixed) {
            A(s->dalloc(&L.Jf, (size_t)B * 26 * L.N));
            A(s->dalloc(&L.diagf, (size_t)B * L.n));
            auto padded_f = [&](float **p) {   // like `padded`: a spare (halo) row on both sides
                float *q = nullptr;
                int r2 = s->dalloc(&q, (size_t)B * L.n + 2 * lrow);
                if (r2 == LDC_OK) {
                    cudaMemset(q, 0, sizeof(float) * ((size_t)B * L.n + 2 * lrow));
                    *p = q + lrow;
                }
                return r2;
            };
            A(padded_f(&L.xf));
            A(padded_f(&L.bf));
            A(padded_f(&L.rf));
            A(s->dalloc(&L.dpf, (size_t)B * L.N));
        }
    }
    // Krylov workspace
    A(s->dalloc(&s->V, (size_t)(m + 1) * B * n));
    s->zs = (long)B * n + 2 * row;
    {
        double *zq = nullptr;
        A(s->dalloc(&zq, (size_t)m * s->zs));
        if (zq) cudaMemset(zq, 0, sizeof(double) * (size_t)m * s->zs);
        s->Z = zq ? zq + row : nullptr;
    }
    A(s->dalloc(&s->w, (size_t)B * n));
    A(s->dalloc(&s->rres, (size_t)B * n));
    const int nblk = dot_blocks(n);
    s->partial_doubles = (long)B * (m + 1) * nblk;
    A(s->dalloc(&s->partials, (size_t)s->partial_doubles));
    const int64_t rw = ldc_residual_work_doubles(g);
    A(s->dalloc(&s->res_work, (size_t)(rw > 0 ? rw : 1)));
    A(s->dalloc(&s->ks.H, (size_t)B * (m + 1) * m));
    A(s->dalloc(&s->ks.cs, (size_t)B * m));
    A(s->dalloc(&s->ks.sn, (size_t)B * m));
    A(s->dalloc(&s->ks.gv, (size_t)B * (m + 1)));
    A(s->dalloc(&s->ks.yv, (size_t)B * m));
    A(s->dalloc(&s->ks.hcol, (size_t)B * (m + 2)));
    A(s->dalloc(&s->ks.hcol2, (size_t)B * (m + 2)));
    A(s->dalloc(&s->ks.nrm2, B));
    A(s->dalloc(&s->ks.beta, B));
    A(s->dalloc(&s->ks.res, B));
    A(s->dalloc(&s->ks.hnext, B));
    A(s->dalloc(&s->run_dev, B));
    A(s->dalloc(&s->fin_dev, B));
    A(s->dalloc(&s->scal_dev, (size_t)3 * B + 8));   // + room for the shape agreement of set_comm
    A(s->halloc(&s->scal_host, (size_t)3 * B));
    A(s->halloc(&s->res_host, B));
    A(s->halloc(&s->du_host, B));
    A(s->halloc(&s->mask_host, B));
#undef A
    if (rc != LDC_OK) {
        ldc_solver_destroy(s);
        return rc;
    }
    s->dt_host.assign(B, g->dt);
    s->bc_host.assign(B, g->bc);
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMemcpy(s->dt_dev, s->dt_host.data(), sizeof(double) * B, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->bc_dev, s->bc_host.data(), sizeof(double) * B, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(s->X_alloc, 0, sizeof(double) * ((size_t)B * n + 2 * row));
    if (e == cudaSuccess) e = cudaMemset(s->xpad_alloc, 0, sizeof(double) * ((size_t)n + 2 * row));
    if (e != cudaSuccess) {
        ldc_solver_destroy(s);
        return cuda_fail(e, "solver init", __FILE__, __LINE__);
    }
    s->g.dt_dev = s->dt_dev;
    s->g.bc_dev = s->bc_dev;
    *out = s;
    return LDC_OK;
}

extern "C" void ldc_solver_destroy(ldc_solver *s)
{
    if (!s) return;
    for (cudaGraphExec_t ge : s->iter_graph) if (ge) cudaGraphExecDestroy(ge);
    if (s->ev_in) cudaEventDestroy(s->ev_in);
    if (s->stream) cudaStreamDestroy(s->stream);
    for (void *p : s->dev_allocs) cudaFree(p);
    for (void *p : s->host_allocs) cudaFreeHost(p);
    delete s;
}

extern "C" double *ldc_solver_state_dev(ldc_solver *s) { return s ? s->X : nullptr; }

extern "C" int ldc_solver_set_comm(ldc_solver *s, const ldc_comm_ops *ops)
{
    LDC_REQUIRE(s && ops && ops->halo_exchange && ops->allreduce && ops->allgather,
                "ldc_solver_set_comm: incomplete function table");
    LDC_REQUIRE(s->batch == 1, "a distributed solver holds one cavity");
    // The hierarchy (levels, replication level), the Krylov limits derived from it and the
    // all-gather layout (rank r's piece at r*count) were decided from THIS rank's row_begin /
    // row_count.  Ranks that disagree would issue different sequences of collectives and hang, so
    // agree here, collectively: every rank sees the minimum and maximum of each quantity and all
    // of them refuse a partition that is not uniform.
    {
        int l_rep = -1;
        for (size_t l = 0; l < s->levels.size(); ++l)
            if (s->levels[l].replicated) { l_rep = (int)l; break; }
        const double mine[3] = {(double)s->levels[0].ny, (double)s->levels.size(), (double)l_rep};
        double host[6];
        for (int k = 0; k < 3; ++k) { host[2 * k] = mine[k]; host[2 * k + 1] = -mine[k]; }
        LDC_CUDA(cudaMemcpy(s->scal_dev, host, sizeof(host), cudaMemcpyHostToDevice));
        if (int rc = ops->allreduce(ops->ctx, s->scal_dev, 6, 1, nullptr)) {
            ldc::set_error("ldc_solver_set_comm: all-reduce of the strip shapes failed (%d)", rc);
            return LDC_ERR_CUDA;
        }
        LDC_CUDA(cudaStreamSynchronize(nullptr));
        LDC_CUDA(cudaMemcpy(host, s->scal_dev, sizeof(host), cudaMemcpyDeviceToHost));
        static const char *what[3] = {"rows per strip", "multigrid levels", "replication level"};
        for (int k = 0; k < 3; ++k)
            LDC_REQUIRE(host[2 * k] == -host[2 * k + 1],
                        "ldc_solver_set_comm: %s differ across ranks (%g..%g, this rank %g); the row strips "
                        "must be equal (ny divisible by the number of ranks)",
                        what[k], -host[2 * k + 1], host[2 * k], mine[k]);
    }
    s->comm = *ops;
    s->has_comm = true;
    // NCCL point-to-point and all-reduce calls are stream-ordered and can be captured; the
    // environment switch is an escape hatch
    if (getenv("LDC_STRIP_NO_GRAPHS") != nullptr) s->use_graphs = false;
    return LDC_OK;
}

extern "C" int ldc_solver_set_keep_diverged(ldc_solver *s, int32_t keep)
{
    LDC_REQUIRE(s, "ldc_solver_set_keep_diverged: null solver");
    s->keep_diverged = keep != 0;
    return LDC_OK;
}

extern "C" int ldc_solver_num_levels(ldc_solver *s) { return s ? (int)s->levels.size() : 0; }

extern "C" int ldc_solver_counters(ldc_solver *s, int64_t *out3)
{
    LDC_REQUIRE(s && out3, "ldc_solver_counters: null argument");
    out3[0] = s->lockstep_krylov_its;   // Krylov iterations executed (a batch advances in lock step)
    out3[1] = s->total_linear_its;      // sum over cavities of their own Krylov iterations
    out3[2] = s->lockstep_newton_its;   // Newton iterations executed (Jacobian builds)
    return LDC_OK;
}

extern "C" int ldc_solver_set_state(ldc_solver *s, const double *X_dev, void *stream)
{
    LDC_REQUIRE(s && X_dev, "ldc_solver_set_state: null argument");
    LDC_CUDA(cudaMemcpyAsync(s->X, X_dev, sizeof(double) * s->batch * s->n, cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    return LDC_OK;
}

extern "C" int ldc_solver_get_state(ldc_solver *s, double *X_dev, void *stream)
{
    LDC_REQUIRE(s && X_dev, "ldc_solver_get_state: null argument");
    LDC_CUDA(cudaMemcpyAsync(X_dev, s->X, sizeof(double) * s->batch * s->n, cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    return LDC_OK;
}

extern "C" int ldc_solver_set_dt(ldc_solver *s, const double *dt_host)
{
    LDC_REQUIRE(s && dt_host, "ldc_solver_set_dt: null argument");
    for (int b = 0; b < s->batch; ++b) {
        LDC_REQUIRE(dt_host[b] > 0.0, "dt[%d] must be positive", b);
        s->dt_host[b] = dt_host[b];
    }
    LDC_CUDA(cudaMemcpy(s->dt_dev, s->dt_host.data(), sizeof(double) * s->batch, cudaMemcpyHostToDevice));
    return LDC_OK;
}

extern "C" int ldc_solver_set_bc(ldc_solver *s, const double *bc_host)
{
    LDC_REQUIRE(s && bc_host, "ldc_solver_set_bc: null argument");
    for (int b = 0; b < s->batch; ++b) s->bc_host[b] = bc_host[b];
    LDC_CUDA(cudaMemcpy(s->bc_dev, s->bc_host.data(), sizeof(double) * s->batch, cudaMemcpyHostToDevice));
    return LDC_OK;
}

// --------------------------------------------------------------------------------------------
static int upload_mask(ldc_solver *s, const std::vector<int32_t> &mask, int32_t *dst, cudaStream_t st)
{
    // mask_host is reused: make sure the previous upload has been consumed
    LDC_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < s->batch; ++b) s->mask_host[b] = mask[b];
    LDC_CUDA(cudaMemcpyAsync(dst, s->mask_host, sizeof(int32_t) * s->batch, cudaMemcpyHostToDevice, st));
    return LDC_OK;
}

static bool any(const std::vector<int32_t> &m)
{
    for (int32_t v : m) if (v) return true;
    return false;
}

// ---- row-strip helpers -----------------------------------------------------------------------
static int comm_allreduce(ldc_solver *s, double *buf, long count, int op, cudaStream_t st)
{
    if (!s->has_comm || s->comm.nranks <= 1) return LDC_OK;
    if (s->comm.allreduce(s->comm.ctx, buf, count, op, st) != 0) {
        set_error("comm all-reduce failed");
        return LDC_ERR_STATE;
    }
    return LDC_OK;
}

static int comm_halo(ldc_solver *s, double *first_owned_row, cudaStream_t st)
{
    if (!s->has_comm || s->comm.nranks <= 1) return LDC_OK;
    if (s->comm.halo_exchange(s->comm.ctx, first_owned_row, 3L * s->g.nx, s->g.row_count, st) != 0) {
        set_error("comm halo exchange failed");
        return LDC_ERR_STATE;
    }
    return LDC_OK;
}

static int level_halo(ldc_solver *s, const MgLevel &L, double *first_owned_row, cudaStream_t st)
{
    if (!s->has_comm || s->comm.nranks <= 1 || L.replicated) return LDC_OK;
    if (s->comm.halo_exchange(s->comm.ctx, first_owned_row, 3L * L.nx, L.ny, st) != 0) {
        set_error("comm halo exchange failed");
        return LDC_ERR_STATE;
    }
    return LDC_OK;
}

// this rank's rows of a replicated level, seen as a strip (for the transfer kernels)
static MgLevel slice_view(const MgLevel &L)
{
    MgLevel v = L;
    const long off = 3L * L.nx * L.slice_begin;
    v.ny = L.slice_rows;
    v.row_begin = L.slice_begin;
    v.N = (long)L.nx * L.slice_rows;
    v.n = 3 * v.N;
    v.ghost_lo = L.slice_begin > 0;
    v.ghost_hi = L.slice_begin + L.slice_rows < L.ny_global;
    v.v_rows = (v.row_begin + v.ny < L.ny_global) ? v.ny : L.ny_global - 1 - v.row_begin;
    if (v.v_rows < 0) v.v_rows = 0;
    v.X = L.X + off;
    v.x = L.x ? L.x + off : nullptr;
    v.b = L.b ? L.b + off : nullptr;
    v.xf = L.xf ? L.xf + off : nullptr;
    v.bf = L.bf ? L.bf + off : nullptr;
    return v;
}

// every rank has filled its slice of a replicated level vector: make it whole everywhere
static int gather_level(ldc_solver *s, const MgLevel &L, double *vec, cudaStream_t st)
{
    if (!s->has_comm || s->comm.nranks <= 1) return LDC_OK;
    if (s->comm.allgather(s->comm.ctx, vec, 3L * L.nx * L.slice_rows, st) != 0) {
        set_error("comm all-gather failed");
        return LDC_ERR_STATE;
    }
    return LDC_OK;
}

// typed views of a level's arrays
template <typename T> struct Lv;
template <> struct Lv<double> {
    static const double *J(const MgLevel &l) { return l.J; }
    static const double *idiag(const MgLevel &l) { return l.diag; }
    static double *x(const MgLevel &l) { return l.x; }
    static double *b(const MgLevel &l) { return l.b; }
    static double *r(const MgLevel &l) { return l.r; }
    static double *dp(const MgLevel &l) { return l.dp; }
};
template <> struct Lv<float> {
    static const float *J(const MgLevel &l) { return l.Jf; }
    static const float *idiag(const MgLevel &l) { return l.diagf; }
    static float *x(const MgLevel &l) { return l.xf; }
    static float *b(const MgLevel &l) { return l.bf; }
    static float *r(const MgLevel &l) { return l.rf; }
    static float *dp(const MgLevel &l) { return l.dpf; }
};

static int level_halo_t(ldc_solver *s, const MgLevel &L, double *v, cudaStream_t st) { return level_halo(s, L, v, st); }
// fp32 level vector of a strip: its rows of 3*nx floats travel as 3*nx/2 doubles (nx is even on
// every distributed level of a mixed-precision solver, see ldc_solver_create)
static int level_halo_t(ldc_solver *s, const MgLevel &L, float *v, cudaStream_t st)
{
    if (!s->has_comm || s->comm.nranks <= 1 || L.replicated) return LDC_OK;
    if (s->comm.halo_exchange(s->comm.ctx, reinterpret_cast<double *>(v), 3L * L.nx / 2, L.ny, st) != 0) {
        set_error("comm halo exchange failed");
        return LDC_ERR_STATE;
    }
    return LDC_OK;
}
static int gather_level_t(ldc_solver *s, const MgLevel &L, double *vec, cudaStream_t st) { return gather_level(s, L, vec, st); }
static int gather_level_t(ldc_solver *s, const MgLevel &L, float *vec, cudaStream_t st)
{
    if (!s->has_comm || s->comm.nranks <= 1) return LDC_OK;
    if (s->comm.allgather(s->comm.ctx, reinterpret_cast<double *>(vec), 3L * L.nx * L.slice_rows / 2, st) != 0) {
        set_error("comm all-gather failed");
        return LDC_ERR_STATE;
    }
    return LDC_OK;
}

// y = J_l x or b - J_l x with the true rows of level l; x must be a padded vector (its halo rows
// are refreshed here for a strip)
template <typename T>
static int level_spmv(ldc_solver *s, const MgLevel &L, T *x_padded, const T *b, T *y, const int32_t *run,
                      cudaStream_t st)
{
    TRY(level_halo_t(s, L, x_padded, st));
    return spmv_compact_launch<T>(L.nx, L.ny, s->batch, Lv<T>::J(L), x_padded, b, y, run, st, L.ghost_lo, L.ghost_hi);
}

// same for an unpadded x (copied into the padded scratch first; strips only)
static int spmv_outer(ldc_solver *s, const double *x, const double *b, double *y, cudaStream_t st)
{
    const MgLevel &L0 = s->levels[0];
    if (!s->strip)
        return spmv_compact_launch<double>(L0.nx, L0.ny, s->batch, L0.J, x, b, y, s->run_dev, st);
    TRY(vec_copy(x, s->xpad, s->n, 1, nullptr, st));
    return level_spmv<double>(s, L0, s->xpad, b, y, s->run_dev, st);
}

// Jacobian at the current state on every level + diagonals
static int build_hierarchy(ldc_solver *s, cudaStream_t st)
{
    for (size_t l = 0; l < s->levels.size(); ++l) {
        MgLevel &L = s->levels[l];
        if (l > 0) {
            const MgLevel &Fl = s->levels[l - 1];
            if (L.replicated && !Fl.replicated) {   // strip -> replicated: restrict own rows, gather
                const MgLevel view = slice_view(L);
                TRY(mg_restrict_state(Fl, view, s->batch, st));
                TRY(gather_level(s, L, L.X, st));
            } else {
                TRY(mg_restrict_state(Fl, L, s->batch, st));
            }
        }
        TRY(level_halo(s, L, L.X, st));   // a strip's Jacobian rows read the neighbours' boundary rows
        ldc_grid gl = level_grid(s, L);
        TRY(fd_jacobian_launch(&gl, L.X, s->U_old, s->V_old, nullptr, L.J, st));
        if (s->opt.pc_type != LDC_PC_NONE) TRY(mg_extract_diag(L, s->batch, st));
    }
    return LDC_OK;
}

// one relaxation sweep x += omega B^-1 (b - J x); the residual uses the true operator (halo of x),
// the cell-block solves are local to the strip
template <typename T>
static int sweep(ldc_solver *s, const MgLevel &L, const T *b, T *x, bool zero_guess, const int32_t *run,
                 cudaStream_t st)
{
    const T *r = b;
    if (!zero_guess) {
        TRY(level_spmv<T>(s, L, x, b, Lv<T>::r(L), run, st));
        r = Lv<T>::r(L);
    }
    return mg_vanka_relax<T>(L, r, Lv<T>::idiag(L), Lv<T>::dp(L), x, zero_guess, s->opt.pc_omega_u,
                             s->opt.pc_omega_p, s->batch, run, st);
}

// x_P -= mean(x_P) over the WHOLE level (all strips)
template <typename T>
static int remove_mean(ldc_solver *s, const MgLevel &L, T *x, const int32_t *run, cudaStream_t st)
{
    if (!s->strip || L.replicated)
        return mg_remove_mean_pressure<T>(L, x, s->batch, run, s->partials, s->ks.hcol2, st);
    TRY(mg_pressure_sum<T>(L, x, 1, run, s->partials, s->ks.hcol2, st));
    TRY(comm_allreduce(s, s->ks.hcol2, 1, 0, st));
    return mg_subtract_mean<T>(L, x, 1, run, s->ks.hcol2, (long)L.nx * L.ny_global, st);
}

// x must be a padded vector of level l
template <typename T>
static int vcycle(ldc_solver *s, size_t l, const T *b, T *x, const int32_t *run, cudaStream_t st)
{
    const MgLevel &L = s->levels[l];
    if (l + 1 == s->levels.size()) {
        // the coarsest grid is relaxed until it is (roughly) solved: a few sweeps on a 4x4 grid,
        // ~2 per cell row when the grid could not be coarsened further (e.g. 100 -> 50 -> 25)
        int sweeps = s->opt.pc_sweeps;
        if (s->levels.size() > 1) {
            sweeps = s->opt.mg_coarse_sweeps;
            const int side = L.nx > L.ny_global ? L.nx : L.ny_global;
            if (sweeps < 2 * side) sweeps = 2 * side;
        }
        for (int k = 0; k < sweeps; ++k) TRY(sweep<T>(s, L, b, x, k == 0, run, st));
        if (s->levels.size() > 1) TRY(remove_mean<T>(s, L, x, run, st));
        return LDC_OK;
    }
    for (int k = 0; k < s->opt.pc_sweeps; ++k) TRY(sweep<T>(s, L, b, x, k == 0, run, st));
    T *r = Lv<T>::r(L);
    TRY(level_spmv<T>(s, L, x, b, r, run, st));
    const MgLevel &C = s->levels[l + 1];
    TRY(level_halo_t(s, L, r, st));   // the last coarse V row of a strip reads the fine row above it
    if (C.replicated && !L.replicated) {
        // distributed -> replicated: every rank restricts into its rows of the coarse
        // right-hand side, one all-gather makes it whole, the rest of the cycle runs without
        // communication, and each rank interpolates from its own rows (+ the row below)
        const MgLevel view = slice_view(C);
        TRY(mg_restrict_residual<T>(L, r, view, Lv<T>::b(view), s->batch, run, st));
        TRY(gather_level_t(s, C, Lv<T>::b(C), st));
        TRY(vcycle<T>(s, l + 1, Lv<T>::b(C), Lv<T>::x(C), run, st));
        TRY(mg_prolong_add<T>(view, Lv<T>::x(view), L, x, s->batch, run, st));
    } else {
        TRY(mg_restrict_residual<T>(L, r, C, Lv<T>::b(C), s->batch, run, st));
        TRY(vcycle<T>(s, l + 1, Lv<T>::b(C), Lv<T>::x(C), run, st));
        TRY(level_halo_t(s, C, Lv<T>::x(C), st));   // the first fine V row of a strip interpolates from below
        TRY(mg_prolong_add<T>(C, Lv<T>::x(C), L, x, s->batch, run, st));
    }
    for (int k = 0; k < s->opt.pc_sweeps; ++k) TRY(sweep<T>(s, L, b, x, false, run, st));
    return LDC_OK;
}

// z = M v ; z is a padded basis vector Z_k
static int apply_pc(ldc_solver *s, const double *v, double *z, const int32_t *run, cudaStream_t st)
{
    if (s->opt.pc_type == LDC_PC_NONE) return vec_copy(v, z, s->n, s->batch, run, st);
    // The Jacobian is singular by the constant-pressure mode (with finite-difference
    // noise: nearly singular).  Searching only among zero-mean pressure corrections keeps the
    // Krylov solution out of that mode (what KSPSetNullSpace does in PETSc).
    if (s->mixed) {
        // the cycle is only a preconditioner of the flexible GMRES: run it in fp32 (half the bytes
        // on its bandwidth-bound levels); the Krylov vectors, the operator and every norm stay fp64
        const MgLevel &L0 = s->levels[0];
        const long n = (long)s->batch * s->n;
        TRY((vec_convert<double, float>(v, L0.bf, n, st)));
        TRY(vcycle<float>(s, 0, L0.bf, L0.xf, run, st));
        TRY(remove_mean<float>(s, L0, L0.xf, run, st));
        return vec_convert<float, double>(L0.xf, z, n, st);
    }
    TRY(vcycle<double>(s, 0, v, z, run, st));
    return remove_mean<double>(s, s->levels[0], z, run, st);
}

// squared norm of x per instance -> dst_dev[b]
static int norm2_to(ldc_solver *s, const double *x, double *dst_dev, const int32_t *run, cudaStream_t st)
{
    const int nblk = dot_blocks(s->n);
    TRY(vec_norm2_partials(x, s->n, s->batch, run, s->partials, st));
    TRY(vec_reduce_partials(s->partials, 1, nblk, s->batch, dst_dev, st));
    return comm_allreduce(s, dst_dev, s->batch, 0, st);
}

// every launch of Krylov iteration k: z_k = M v_k, w = J z_k, Gram-Schmidt against v_0..v_k,
// Givens update of the Hessenberg column, v_{k+1} = w / h_{k+1,k}
static int krylov_iteration_body(ldc_solver *s, int k, cudaStream_t st)
{
    const int B = s->batch, m = s->m;
    const long n = s->n;
    const long vs = (long)B * n;
    const int nblk = dot_blocks(n);
    const MgLevel &L0 = s->levels[0];
    const ldc_solver_options &o = s->opt;
    const int tb = 64, gb = (B + tb - 1) / tb;
    double *vk = s->V + (long)k * vs, *zk = s->Z + (long)k * s->zs;
    TRY(apply_pc(s, vk, zk, s->run_dev, st));
    TRY(level_spmv<double>(s, L0, zk, nullptr, s->w, s->run_dev, st));
    TRY(vec_multi_dot(s->V, vs, k + 1, s->w, n, B, s->run_dev, s->partials, st));
    TRY(vec_reduce_partials(s->partials, k + 1, nblk, B, s->ks.hcol, st));
    TRY(comm_allreduce(s, s->ks.hcol, (long)B * (k + 1), 0, st));
    TRY(vec_gs_update(s->V, vs, k + 1, s->ks.hcol, k + 1, s->w, n, B, s->run_dev, s->partials, st));
    if (o.gmres_refine) {
        // second classical Gram-Schmidt pass (CGS2): keeps the basis orthogonal to working
        // precision so the residual estimate stays equal to the true residual
        TRY(vec_multi_dot(s->V, vs, k + 1, s->w, n, B, s->run_dev, s->partials, st));
        TRY(vec_reduce_partials(s->partials, k + 1, nblk, B, s->ks.hcol2, st));
        TRY(comm_allreduce(s, s->ks.hcol2, (long)B * (k + 1), 0, st));
        TRY(vec_gs_update(s->V, vs, k + 1, s->ks.hcol2, k + 1, s->w, n, B, s->run_dev, s->partials, st));
        add_columns_kernel<<<(B * (k + 1) + 127) / 128, 128, 0, st>>>(s->ks.hcol, s->ks.hcol2, B * (k + 1));
        LDC_CHECK_LAUNCH();
    }
    TRY(vec_reduce_partials(s->partials, 1, nblk, B, s->ks.nrm2, st));
    TRY(comm_allreduce(s, s->ks.nrm2, B, 0, st));
    krylov_givens_kernel<<<gb, tb, 0, st>>>(s->ks, m, k, B, s->run_dev);
    LDC_CHECK_LAUNCH();
    TRY(vec_scale_dev(s->w, s->ks.hnext, 1, true, s->V + (long)(k + 1) * vs, n, B, s->run_dev, st));
    return LDC_OK;
}

static int krylov_iteration(ldc_solver *s, int k, cudaStream_t st)
{
    if (!s->use_graphs) return krylov_iteration_body(s, k, st);
    if (!s->iter_graph[k]) {
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            s->use_graphs = false;
            return krylov_iteration_body(s, k, st);
        }
        const int rc = krylov_iteration_body(s, k, st);
        const cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (rc != LDC_OK || e != cudaSuccess || graph == nullptr ||
            cudaGraphInstantiate(&s->iter_graph[k], graph, 0) != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            s->iter_graph[k] = nullptr;
            s->use_graphs = false;
            return krylov_iteration_body(s, k, st);
        }
        cudaGraphDestroy(graph);
    }
    LDC_CUDA(cudaGraphLaunch(s->iter_graph[k], st));
    return LDC_OK;
}

// Solve J y = rhs for the instances flagged in `run` (host vector, modified).  bnorm = ||rhs||.
static int fgmres(ldc_solver *s, const double *rhs, double *y, std::vector<int32_t> run,
                  const std::vector<double> &bnorm, std::vector<int32_t> &reason,
                  std::vector<int32_t> &its, std::vector<double> &relres, cudaStream_t st)
{
    const int B = s->batch, m = s->m;
    const long n = s->n;
    const long vs = (long)B * n;  // stride between basis vectors
    const int nblk = dot_blocks(n);
    const MgLevel &L0 = s->levels[0];
    const ldc_solver_options &o = s->opt;
    const int tb = 64, gb = (B + tb - 1) / tb;

    // A grid that cannot be coarsened (odd size) only has the single-level relaxation as
    // preconditioner, whose iteration count grows with the grid; the fail-fast limit tuned for the
    // multigrid cycle would turn every such solve into a spurious "diverged".
    int ksp_max_it = o.ksp_max_it;
    if (s->levels.size() == 1 && o.pc_type != LDC_PC_NONE) {
        const int single = 10 * (s->g.nx + s->g.ny);
        if (ksp_max_it < single) ksp_max_it = single;
    }
    reason.assign(B, 0);
    its.assign(B, 0);
    relres.assign(B, 0.0);
    std::vector<double> cycle_start(B, 1.0);   // relative residual at the start of the restart cycle
    static const bool debug_ksp = getenv("LDC_DEBUG_KSP") != nullptr;   // not on the per-iteration path
    if (debug_ksp) fprintf(stderr, "[ksp] new solve bnorm=%.3e\n", bnorm[0]);
    // y = 0 for EVERY requested instance, also those that finish right here with a zero right-hand
    // side (reason 3): their answer is the zero vector, not whatever the caller's buffer held
    TRY(upload_mask(s, run, s->run_dev, st));
    TRY(vec_zero(y, n, B, s->run_dev, st));
    {
        bool dropped = false;
        for (int b = 0; b < B; ++b)
            if (run[b] && !(bnorm[b] > 0.0)) { run[b] = 0; reason[b] = (bnorm[b] == 0.0) ? 3 : -9; dropped = true; }
        if (dropped) TRY(upload_mask(s, run, s->run_dev, st));
    }
    // r0 = rhs, beta = ||rhs||
    TRY(vec_copy(rhs, s->rres, n, B, s->run_dev, st));
    TRY(norm2_to(s, s->rres, s->ks.nrm2, s->run_dev, st));
    sqrt_kernel<<<gb, tb, 0, st>>>(s->ks.nrm2, s->ks.beta, B, s->run_dev);
    LDC_CHECK_LAUNCH();

    while (any(run)) {
        TRY(vec_scale_dev(s->rres, s->ks.beta, 1, true, s->V, n, B, s->run_dev, st));
        krylov_begin_cycle_kernel<<<gb, tb, 0, st>>>(s->ks, m, B, s->run_dev);
        LDC_CHECK_LAUNCH();
        int k = 0;
        for (; k < m && any(run); ++k) {
            TRY(krylov_iteration(s, k, st));
            s->lockstep_krylov_its += 1;
            LDC_CUDA(cudaMemcpyAsync(s->res_host, s->ks.res, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
            LDC_CUDA(cudaStreamSynchronize(st));
            std::vector<int32_t> fin(B, 0);
            bool any_fin = false;
            for (int b = 0; b < B; ++b) {
                if (!run[b]) continue;
                its[b] += 1;
                const double r = s->res_host[b];
                relres[b] = r / bnorm[b];
                int rs = 0;
                if (r != r) rs = -9;
                else if (r <= o.ksp_atol) rs = 3;
                else if (r <= o.ksp_rtol * bnorm[b]) rs = 2;
                else if (r >= o.ksp_dtol * bnorm[b]) rs = -4;
                else if (its[b] >= ksp_max_it) rs = -3;
                if (rs != 0) { reason[b] = rs; fin[b] = 1; any_fin = true; }
            }
            if (any_fin) {
                TRY(upload_mask(s, fin, s->fin_dev, st));
                krylov_solve_y_kernel<<<gb, tb, 0, st>>>(s->ks, m, k + 1, B, s->fin_dev);
                LDC_CHECK_LAUNCH();
                TRY(vec_multi_axpy(s->Z, s->zs, k + 1, s->ks.yv, m, y, n, B, s->fin_dev, st));
                for (int b = 0; b < B; ++b) if (fin[b]) run[b] = 0;
                TRY(upload_mask(s, run, s->run_dev, st));
            }
        }
        if (debug_ksp) fprintf(stderr, "[ksp] cycle end: its=%d relres=%.3e\n", its[0], relres[0]);
        // Stagnation: a full restart cycle that removes less than 10% of the residual will not reach
        // the tolerance within the iteration cap (seen when the Newton iterate has blown up: relres
        // 0.92 -> 0.92 -> 0.92 ...); healthy cycles reduce it by 4x or more.  Failing now instead of
        // at ksp_max_it is the same outcome (reason -3 -> Newton diverged -> dt halved), 4 cycles sooner.
        {
            bool changed = false;
            for (int b = 0; b < B; ++b) {
                if (!run[b]) continue;
                if (relres[b] > o.ksp_stagnation * cycle_start[b]) { reason[b] = -3; run[b] = 0; changed = true; }
                cycle_start[b] = relres[b];
            }
            if (changed && any(run)) TRY(upload_mask(s, run, s->run_dev, st));
        }
        if (!any(run)) break;
        // restart: fold the cycle into y, recompute the true residual
        krylov_solve_y_kernel<<<gb, tb, 0, st>>>(s->ks, m, k, B, s->run_dev);
        LDC_CHECK_LAUNCH();
        TRY(vec_multi_axpy(s->Z, s->zs, k, s->ks.yv, m, y, n, B, s->run_dev, st));
        TRY(spmv_outer(s, y, rhs, s->rres, st));
        TRY(norm2_to(s, s->rres, s->ks.nrm2, s->run_dev, st));
        sqrt_kernel<<<gb, tb, 0, st>>>(s->ks.nrm2, s->ks.beta, B, s->run_dev);
        LDC_CHECK_LAUNCH();
    }
    return LDC_OK;
}

static int residual_and_norm(ldc_solver *s, double *norm2_dst, cudaStream_t st)
{
    TRY(comm_halo(s, s->X, st));
    TRY(ldc_residual(&s->g, s->X, s->U_old, s->V_old, s->F, norm2_dst, s->res_work,
                     s->opt.residual_variant, st));
    return norm2_dst ? comm_allreduce(s, norm2_dst, s->batch, 0, st) : LDC_OK;
}

// Building blocks of the (distributed) solve, timed on the solver's own data with CUDA events on
// the solver's stream: exactly the call sequences the Newton / Krylov loops issue, halo exchanges
// and all-reduces included.  bench.py uses it for the row-strip table (BASELINE config 5).
extern "C" int ldc_solver_time_ops(ldc_solver *s, int32_t iters, double *ms_out4, void *stream)
{
    LDC_REQUIRE(s && ms_out4 && iters >= 1, "ldc_solver_time_ops: bad argument");
    LDC_CUDA(cudaEventRecord(s->ev_in, (cudaStream_t)stream));
    cudaStream_t st = s->stream;
    LDC_CUDA(cudaStreamWaitEvent(st, s->ev_in, 0));
    const int B = s->batch;
    const long n = s->n, vs = (long)B * n;
    const int nblk = dot_blocks(n);
    const int kdot = s->m < 15 ? s->m : 15;   // the average Gram-Schmidt width of a GMRES(30) cycle
    std::vector<int32_t> all(B, 1);
    TRY(upload_mask(s, all, s->run_dev, st));
    TRY(build_hierarchy(s, st));              // a valid operator on every level
    // defined inputs: the basis vectors the timed operations read = the current residual
    TRY(residual_and_norm(s, s->scal_dev, st));
    for (int k = 0; k < kdot; ++k) TRY(vec_copy(s->F, s->V + (long)k * vs, n, B, s->run_dev, st));
    TRY(vec_copy(s->F, s->w, n, B, s->run_dev, st));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    LDC_CUDA(cudaEventCreate(&e0));
    LDC_CUDA(cudaEventCreate(&e1));
    int rc = LDC_OK;
    for (int op = 0; op < 4 && rc == LDC_OK; ++op) {
        for (int pass = 0; pass < 2 && rc == LDC_OK; ++pass) {   // pass 0 = warm-up
            const int reps = pass == 0 ? 2 : iters;
            if (pass == 1) cudaEventRecord(e0, st);
            for (int i = 0; i < reps && rc == LDC_OK; ++i) {
                switch (op) {
                    case 0:   // residual + its norm: halo of X, kernel, all-reduce of sum(F^2)
                        rc = residual_and_norm(s, s->scal_dev, st);
                        break;
                    case 1:   // outer operator: w = J v (halo of v inside)
                        rc = spmv_outer(s, s->V + (long)(i % 2) * vs, nullptr, s->w, st);
                        break;
                    case 2:   // classical Gram-Schmidt projection against kdot basis vectors: fused dots,
                              // all-reduce of the coefficients, update
                        rc = vec_multi_dot(s->V, vs, kdot, s->w, n, B, s->run_dev, s->partials, st);
                        if (rc == LDC_OK) rc = vec_reduce_partials(s->partials, kdot, nblk, B, s->ks.hcol, st);
                        if (rc == LDC_OK) rc = comm_allreduce(s, s->ks.hcol, (long)B * kdot, 0, st);
                        break;
                    default:  // one application of the preconditioner (multigrid cycle)
                        rc = apply_pc(s, s->V, s->Z, s->run_dev, st);
                        break;
                }
            }
            if (pass == 1) cudaEventRecord(e1, st);
        }
        if (rc == LDC_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = LDC_ERR_CUDA;
        float ms = 0.f;
        if (rc == LDC_OK) cudaEventElapsedTime(&ms, e0, e1);
        ms_out4[op] = (double)ms / iters;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc == LDC_ERR_CUDA) return ldc::cuda_fail(cudaGetLastError(), "ldc_solver_time_ops", __FILE__, __LINE__);
    return rc;
}

extern "C" int ldc_solver_linear_solve(ldc_solver *s, const double *rhs_dev, double *y_dev,
                                       int32_t *its_host, int32_t *reason_host, double *relres_host,
                                       void *stream)
{
    LDC_REQUIRE(s && rhs_dev && y_dev, "ldc_solver_linear_solve: null argument");
    LDC_CUDA(cudaEventRecord(s->ev_in, (cudaStream_t)stream));
    cudaStream_t st = s->stream;
    LDC_CUDA(cudaStreamWaitEvent(st, s->ev_in, 0));
    const int B = s->batch;
    std::vector<int32_t> all(B, 1);
    TRY(upload_mask(s, all, s->run_dev, st));
    extract_old_kernel<<<dim3((unsigned)((s->N + 255) / 256), B), 256, 0, st>>>(
        s->g.nx, s->g.row_count, s->v_rows, s->X, s->U_old, s->V_old, s->run_dev);
    LDC_CHECK_LAUNCH();
    TRY(build_hierarchy(s, st));
    TRY(norm2_to(s, rhs_dev, s->scal_dev, nullptr, st));
    LDC_CUDA(cudaMemcpyAsync(s->scal_host, s->scal_dev, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    LDC_CUDA(cudaStreamSynchronize(st));
    std::vector<double> bnorm(B);
    for (int b = 0; b < B; ++b) bnorm[b] = sqrt(s->scal_host[b]);
    std::vector<int32_t> reason, its;
    std::vector<double> relres;
    TRY(fgmres(s, rhs_dev, y_dev, all, bnorm, reason, its, relres, st));
    LDC_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) {
        if (its_host) its_host[b] = its[b];
        if (reason_host) reason_host[b] = reason[b];
        if (relres_host) relres_host[b] = relres[b];
    }
    return LDC_OK;
}

extern "C" int ldc_solver_step(ldc_solver *s, const int32_t *active_host, ldc_newton_report *rep,
                               double *du_inf_host, void *stream)
{
    LDC_REQUIRE(s && rep, "ldc_solver_step: null argument");
    // work on the private stream, ordered after whatever the caller queued on `stream`; the call
    // synchronises before returning, so later work of the caller is ordered after it
    LDC_CUDA(cudaEventRecord(s->ev_in, (cudaStream_t)stream));
    cudaStream_t st = s->stream;
    LDC_CUDA(cudaStreamWaitEvent(st, s->ev_in, 0));
    const int B = s->batch;
    const long n = s->n;
    const ldc_solver_options &o = s->opt;
    std::vector<int32_t> act(B, 1);
    if (active_host) for (int b = 0; b < B; ++b) act[b] = active_host[b] ? 1 : 0;
    for (int b = 0; b < B; ++b) memset(&rep[b], 0, sizeof(ldc_newton_report));

    // phi_old <- phi for the active cavities (petsc_solver_wrapper.py:53-56)
    TRY(upload_mask(s, act, s->run_dev, st));
    TRY(vec_copy(s->X, s->X_prev, n, B, s->run_dev, st));
    extract_old_kernel<<<dim3((unsigned)((s->N + 255) / 256), B), 256, 0, st>>>(
        s->g.nx, s->g.row_count, s->v_rows, s->X, s->U_old, s->V_old, s->run_dev);
    LDC_CHECK_LAUNCH();

    TRY(residual_and_norm(s, s->scal_dev, st));
    LDC_CUDA(cudaMemcpyAsync(s->scal_host, s->scal_dev, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    LDC_CUDA(cudaStreamSynchronize(st));
    std::vector<int32_t> run = act;
    std::vector<double> fnorm(B, 0.0), ttol(B, 0.0);
    for (int b = 0; b < B; ++b) {
        if (!run[b]) continue;
        fnorm[b] = sqrt(s->scal_host[b]);
        rep[b].fnorm0 = rep[b].fnorm = fnorm[b];
        rep[b].function_evaluations = 1;   // residual kernel launches this instance took part in
        rep[b].jacobian_evaluations = 0;   // FD Jacobian launches (one launch covers all 21 colours)
        ttol[b] = fnorm[b] * o.snes_rtol;
        if (fnorm[b] != fnorm[b]) { rep[b].reason = -4; run[b] = 0; }
        else if (fnorm[b] < o.snes_atol) { rep[b].reason = 2; run[b] = 0; }
    }

    for (int it = 0; it < o.snes_max_it && any(run); ++it) {
        TRY(build_hierarchy(s, st));
        s->lockstep_newton_its += 1;
        std::vector<int32_t> kreason, kits;
        std::vector<double> krel;
        TRY(fgmres(s, s->F, s->Y, run, fnorm, kreason, kits, krel, st));
        for (int b = 0; b < B; ++b) {
            if (!run[b]) continue;
            rep[b].jacobian_evaluations += 1;
            rep[b].linear_iterations += kits[b];
            rep[b].ksp_reason = kreason[b];
            s->total_linear_its += kits[b];
            if (kreason[b] < 0) { rep[b].reason = -3; rep[b].iterations = it; run[b] = 0; }
        }
        if (!any(run)) break;
        TRY(upload_mask(s, run, s->run_dev, st));
        // full step x <- x - y (snes_linesearch_type basic)
        TRY(vec_axpby(1.0, s->X, -1.0, s->Y, s->X, n, B, s->run_dev, st));
        TRY(residual_and_norm(s, s->scal_dev, st));
        TRY(norm2_to(s, s->Y, s->scal_dev + B, s->run_dev, st));
        TRY(norm2_to(s, s->X, s->scal_dev + 2 * B, s->run_dev, st));
        LDC_CUDA(cudaMemcpyAsync(s->scal_host, s->scal_dev, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost, st));
        LDC_CUDA(cudaStreamSynchronize(st));
        for (int b = 0; b < B; ++b) {
            if (!run[b]) continue;
            fnorm[b] = sqrt(s->scal_host[b]);
            const double ynorm = sqrt(s->scal_host[B + b]), xnorm = sqrt(s->scal_host[2 * B + b]);
            rep[b].fnorm = fnorm[b];
            rep[b].ynorm = ynorm;
            rep[b].xnorm = xnorm;
            rep[b].iterations = it + 1;
            rep[b].function_evaluations += 1;
            int rs = 0;
            if (fnorm[b] != fnorm[b]) rs = -4;
            else if (fnorm[b] < o.snes_atol) rs = 2;
            else if (rep[b].function_evaluations >= o.snes_max_funcs) rs = -2;
            else if (fnorm[b] <= ttol[b]) rs = 3;
            else if (ynorm < o.snes_stol * xnorm) rs = 4;
            else if (o.snes_divtol > 0.0 && fnorm[b] > o.snes_divtol * rep[b].fnorm0) rs = -9;
            else if (it + 1 >= o.snes_max_it) rs = -5;
            if (rs != 0) { rep[b].reason = rs; run[b] = 0; }
        }
    }
    for (int b = 0; b < B; ++b)
        if (act[b] && rep[b].reason == 0) rep[b].reason = -5;

    // diverged cavities go back to the state they entered with (the reference raises and the
    // time stepper retries from the old graph, time_stepper.py:23-32)
    std::vector<int32_t> bad(B, 0);
    for (int b = 0; b < B; ++b) bad[b] = (act[b] && rep[b].reason <= 0) ? 1 : 0;
    if (any(bad) && !s->keep_diverged) {
        TRY(upload_mask(s, bad, s->fin_dev, st));
        TRY(vec_copy(s->X_prev, s->X, n, B, s->fin_dev, st));
    }
    if (du_inf_host) {
        TRY(vec_du_inf(s->X, s->U_old, s->g.nx, s->g.row_count, B, s->ks.nrm2, st));
        TRY(comm_allreduce(s, s->ks.nrm2, B, 1, st));
        LDC_CUDA(cudaMemcpyAsync(s->du_host, s->ks.nrm2, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    }
    LDC_CUDA(cudaStreamSynchronize(st));
    if (du_inf_host) for (int b = 0; b < B; ++b) du_inf_host[b] = s->du_host[b];
    return LDC_OK;
}
