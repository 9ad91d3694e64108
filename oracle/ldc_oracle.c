/*
 * ldc_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the lid-driven-cavity hot path of
 * tarcisiofischer/lid_driven_cavity_problem.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (lid_driven_cavity_problem_b200/csrc) never links or calls it.
 *
 * Parity status
 *   residual, X packing, Jacobian mask : PINNED against the reference itself
 *       (oracle/_ref = reference C++ compiled unchanged; reference Python imported
 *       in the build container; reference golden files) -- see tests/test_oracle.py.
 *   coloured-FD Jacobian / ILU(0) / GMRES / Newton : PARITY UNPINNED.  These live in
 *       PETSc (third-party, not vendored, version unpinned in
 *       environment.devenv.yml:12-13) which cannot be installed here.  They restate
 *       PETSc's documented algorithms as selected by
 *       nonlinear_solver/petsc_solver_wrapper.py:112-145.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off [-fopenmp]).
 * -ffp-contract=off keeps the arithmetic identical to the reference's x86-64 -O2
 * build (no fused multiply-add).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LDC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* 1. Residual.  Follows c++/residual_function.cpp:10-238 (serial) and        */
/*    c++/residual_function_omp.cpp:61,90,161 (the three parallel loops);     */
/*    use_cds follows numba_residual_function.py:128-137,202-211.             */
/*    Layout X[3c+0]=P, X[3c+1]=U (dummy when i==nx-1), X[3c+2]=V (dummy when */
/*    j==ny-1), c=i+nx*j  (pure_python_residual_function.py:10-22).           */
/* ------------------------------------------------------------------------- */

typedef struct {
    const double *X;
    long nx, ny;
    double bc;
} field_t;

/* raw fetches on the interleaved vector */
static inline double fP(const field_t *f, long i, long j) { return f->X[3 * (i + f->nx * j)]; }
static inline double fU(const field_t *f, long i, long j) { return f->X[3 * (i + f->nx * j) + 1]; }
static inline double fV(const field_t *f, long i, long j) { return f->X[3 * (i + f->nx * j) + 2]; }

static inline double upwind_beta(double face_velocity, int use_cds)
{
    /* residual_function.cpp:134-137: strict '>' so 0.0 -> -0.5 */
    if (use_cds) return 0.0;
    return face_velocity > 0.0 ? 0.5 : -0.5;
}

/* continuity of cell (i,j): residual_function.cpp:61-86 */
static inline double continuity_row(const field_t *f, long i, long j, double dx, double dy)
{
    const long nx = f->nx, ny = f->ny;
    double U_w = (i == 0) ? 0.0 : fU(f, i - 1, j);
    double U_e = (i == nx - 1) ? 0.0 : fU(f, i, j);
    double V_n = (j == ny - 1) ? 0.0 : fV(f, i, j);
    double V_s = (j == 0) ? 0.0 : fV(f, i, j - 1);
    return (U_e * dy - U_w * dy) + (V_n * dx - V_s * dx);
}

/* x-momentum of U[i,j], 0<=i<=nx-2 : residual_function.cpp:89-156 */
static inline double xmom_row(const field_t *f, long i, long j, double U_P_old, double dt,
                              double dx, double dy, double rho, double mi, int use_cds)
{
    const long nx = f->nx, ny = f->ny;
    const int left = (i == 0), right = (i == nx - 2), bottom = (j == 0), top = (j == ny - 1);

    double U_P = fU(f, i, j);
    double U_W = left ? 0.0 : fU(f, i - 1, j);
    double U_E = right ? 0.0 : fU(f, i + 1, j);
    double U_N = top ? f->bc : fU(f, i, j + 1);
    double U_S = bottom ? 0.0 : fU(f, i, j - 1);
    double P_w = fP(f, i, j);
    double P_e = fP(f, i + 1, j);
    double V_NE = top ? 0.0 : fV(f, i + 1, j);
    double V_NW = top ? 0.0 : fV(f, i, j);
    double V_SE = bottom ? 0.0 : fV(f, i + 1, j - 1);
    double V_SW = bottom ? 0.0 : fV(f, i, j - 1);

    double dU_e_dx = (U_E - U_P) / dx;
    double dU_w_dx = (U_P - U_W) / dx;
    double dU_n_dx = (U_N - U_P) / dy;
    double dU_s_dx = (U_P - U_S) / dy;
    double U_e = (U_E + U_P) / 2.0;
    double U_w = (U_P + U_W) / 2.0;
    double V_n = (V_NE + V_NW) / 2.0;
    double V_s = (V_SE + V_SW) / 2.0;
    double b_Ue = upwind_beta(U_e, use_cds);
    double b_Uw = upwind_beta(U_w, use_cds);
    double b_Vn = upwind_beta(V_n, use_cds);
    double b_Vs = upwind_beta(V_s, use_cds);

    double transient_term = (rho * U_P - rho * U_P_old) * (dx * dy / dt);
    double advective_term =
        rho * U_e * ((.5 - b_Ue) * U_E + (.5 + b_Ue) * U_P) * dy -
        rho * U_w * ((.5 - b_Uw) * U_P + (.5 + b_Uw) * U_W) * dy +
        rho * V_n * ((.5 - b_Vn) * U_N + (.5 + b_Vn) * U_P) * dx -
        rho * V_s * ((.5 - b_Vs) * U_P + (.5 + b_Vs) * U_S) * dx;
    double difusive_term =
        mi * dU_e_dx * dy -
        mi * dU_w_dx * dy +
        mi * dU_n_dx * dx -
        mi * dU_s_dx * dx;
    double source_term = -(P_e - P_w) * dy;
    return transient_term + advective_term - difusive_term - source_term;
}

/* y-momentum of V[i,j], 0<=j<=ny-2 : residual_function.cpp:159-226 */
static inline double ymom_row(const field_t *f, long i, long j, double V_P_old, double dt,
                              double dx, double dy, double rho, double mi, int use_cds)
{
    const long nx = f->nx, ny = f->ny;
    const int left = (i == 0), right = (i == nx - 1), bottom = (j == 0), top = (j == ny - 2);

    double V_P = fV(f, i, j);
    double V_E = right ? 0.0 : fV(f, i + 1, j);
    double V_W = left ? 0.0 : fV(f, i - 1, j);
    double V_N = top ? 0.0 : fV(f, i, j + 1);
    double V_S = bottom ? 0.0 : fV(f, i, j - 1);
    double P_n = fP(f, i, j + 1);
    double P_s = fP(f, i, j);
    double U_SE = right ? 0.0 : fU(f, i, j);
    double U_SW = left ? 0.0 : fU(f, i - 1, j);
    /* quirk kept: lid speed on the top V row although U[.,ny-1] exists (cpp:192-193) */
    double U_NE = right ? 0.0 : (top ? f->bc : fU(f, i, j + 1));
    double U_NW = left ? 0.0 : (top ? f->bc : fU(f, i - 1, j + 1));

    double dV_e_dx = (V_E - V_P) / dx;
    double dV_w_dx = (V_P - V_W) / dx;
    double dV_n_dx = (V_N - V_P) / dy;
    double dV_s_dx = (V_P - V_S) / dy;
    double U_e = (U_NE + U_SE) / 2.0;
    double U_w = (U_NW + U_SW) / 2.0;
    double V_n = (V_P + V_N) / 2.0;
    double V_s = (V_S + V_P) / 2.0;
    double b_Ue = upwind_beta(U_e, use_cds);
    double b_Uw = upwind_beta(U_w, use_cds);
    double b_Vn = upwind_beta(V_n, use_cds);
    double b_Vs = upwind_beta(V_s, use_cds);

    double transient_term = (rho * V_P - rho * V_P_old) * (dx * dy / dt);
    double advective_term =
        rho * U_e * ((.5 - b_Ue) * V_E + (.5 + b_Ue) * V_P) * dy -
        rho * U_w * ((.5 - b_Uw) * V_P + (.5 + b_Uw) * V_W) * dy +
        rho * V_n * ((.5 - b_Vn) * V_N + (.5 + b_Vn) * V_P) * dx -
        rho * V_s * ((.5 - b_Vs) * V_P + (.5 + b_Vs) * V_S) * dx;
    double difusive_term =
        mi * dV_e_dx * dy -
        mi * dV_w_dx * dy +
        mi * dV_n_dx * dx -
        mi * dV_s_dx * dx;
    double source_term = -(P_n - P_s) * dx;
    return transient_term + advective_term - difusive_term - source_term;
}

/* all three rows of cell (i,j); dummy rows F=X (residual_function.cpp:228-235) */
static inline void cell_rows(const field_t *f, long i, long j, const double *U_old,
                             const double *V_old, double dt, double dx, double dy, double rho,
                             double mi, int use_cds, double out[3])
{
    const long nx = f->nx, ny = f->ny;
    const long c = i + nx * j;
    out[0] = continuity_row(f, i, j, dx, dy);
    out[1] = (i < nx - 1)
                 ? xmom_row(f, i, j, U_old[i + (nx - 1) * j], dt, dx, dy, rho, mi, use_cds)
                 : f->X[3 * c + 1];
    out[2] = (j < ny - 1) ? ymom_row(f, i, j, V_old[c], dt, dx, dy, rho, mi, use_cds)
                          : f->X[3 * c + 2];
}

LDC_API int ldc_oracle_residual(const double *X, const double *U_old, const double *V_old,
                                double *F, long nx, long ny, double dt, double dx, double dy,
                                double rho, double mi, double bc, int use_cds, int nthreads)
{
    if (nx < 2 || ny < 2) return -1;
    field_t f = {X, nx, ny, bc};
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
    for (long j = 0; j < ny; ++j) {
        for (long i = 0; i < nx; ++i) {
            cell_rows(&f, i, j, U_old, V_old, dt, dx, dy, rho, mi, use_cds,
                      F + 3 * (i + nx * j));
        }
    }
    return 0;
}

LDC_API int ldc_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* 2. Packing.  nonlinear_solver/_common.py:4-18 (_create_X), :21-36          */
/*    (_recover_X).                                                           */
/* ------------------------------------------------------------------------- */
LDC_API void ldc_oracle_create_X(const double *U, const double *V, const double *P, double *X,
                                 long nx, long ny)
{
    for (long j = 0; j < ny; ++j)
        for (long i = 0; i < nx; ++i) {
            long c = i + nx * j;
            X[3 * c] = P[c];
            X[3 * c + 1] = (i < nx - 1) ? U[i + (nx - 1) * j] : 0.0;
            X[3 * c + 2] = (j < ny - 1) ? V[c] : 0.0;
        }
}

LDC_API void ldc_oracle_recover_X(const double *X, double *U, double *V, double *P, long nx,
                                  long ny)
{
    for (long j = 0; j < ny; ++j)
        for (long i = 0; i < nx; ++i) {
            long c = i + nx * j;
            P[c] = X[3 * c];
            if (i < nx - 1) U[i + (nx - 1) * j] = X[3 * c + 1];
            if (j < ny - 1) V[c] = X[3 * c + 2];
        }
}

/* ------------------------------------------------------------------------- */
/* 3. Jacobian mask.  nonlinear_solver/_common.py:39-57: diags(ones(7),       */
/*    [-nx+1,-nx,-1,0,1,nx,nx-1]) (x) ones(dof,dof), canonical CSR, int32.    */
/*    The diagonals are NOT cut at row ends (wrap-around entries stay).       */
/* ------------------------------------------------------------------------- */
static void mask_offsets(long nx, long off[7])
{
    off[0] = -nx; off[1] = -nx + 1; off[2] = -1; off[3] = 0;
    off[4] = 1;   off[5] = nx - 1;  off[6] = nx;
}

LDC_API long ldc_oracle_mask_nnz(long nx, long ny, long dof)
{
    long N = nx * ny, off[7], cells = 0;
    mask_offsets(nx, off);
    for (int k = 0; k < 7; ++k) {
        long a = off[k] < 0 ? -off[k] : off[k];
        if (a < N) cells += N - a;
    }
    return cells * dof * dof;
}

/* indptr has dof*N+1 entries, indices has nnz entries */
LDC_API int ldc_oracle_mask_csr(long nx, long ny, long dof, int32_t *indptr, int32_t *indices)
{
    if (nx < 3) return -1; /* offsets collide: reference TODO test_calculate_jacobian_mask.py:20-22 */
    long N = nx * ny, off[7];
    mask_offsets(nx, off);
    long pos = 0;
    for (long c = 0; c < N; ++c)
        for (long d = 0; d < dof; ++d) {
            indptr[dof * c + d] = (int32_t)pos;
            for (int k = 0; k < 7; ++k) {
                long cc = c + off[k];
                if (cc < 0 || cc >= N) continue;
                for (long e = 0; e < dof; ++e) indices[pos++] = (int32_t)(dof * cc + e);
            }
        }
    indptr[dof * N] = (int32_t)pos;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* 4. Coloured finite-difference Jacobian, PETSc MatFDColoring 'ds' rule      */
/*    (selected by petsc_solver_wrapper.py:132,140).  PARITY UNPINNED.        */
/*    PETSc (MatFDColoringApply_AIJ): dx=x_c; |dx|<umin -> +-umin; dx*=eps;   */
/*    w = x + dx e_c ; J(r,c) = (F(w)-F(x))_r * (1/dx).                       */
/*    Colours: cell colour (i+3j) mod 7, colour = 3*cell_colour+dof (SURVEY   */
/*    App. C.3); mask wrap-around entries never see a true dependency and are */
/*    exactly 0 (what a column-by-column FD gives).                           */
/* ------------------------------------------------------------------------- */
#define LDC_FD_EPS 1.4901161193847656e-08 /* sqrt(DBL_EPSILON), PETSc error_rel default */
#define LDC_FD_UMIN 1.0e-6                /* PETSc umin default */

static inline double fd_step(double x)
{
    double dx = x;
    if (fabs(dx) < LDC_FD_UMIN) dx = (dx >= 0.0) ? LDC_FD_UMIN : -LDC_FD_UMIN;
    dx *= LDC_FD_EPS;
    return dx;
}

/* is the mask entry (row cell cr -> column cell cc) a wrap-around pair? */
static inline int is_wrap_pair(long cr, long cc, long nx)
{
    long ir = cr % nx, ic = cc % nx;
    long di = ic - ir;
    return di > 1 || di < -1; /* a true stencil neighbour is at most one column away */
}

LDC_API int ldc_oracle_fd_jacobian(const double *X, const double *U_old, const double *V_old,
                                   const double *F0, const int32_t *indptr,
                                   const int32_t *indices, double *values, long nx, long ny,
                                   double dt, double dx, double dy, double rho, double mi,
                                   double bc, int use_cds, int nthreads)
{
    const long N = nx * ny, n = 3 * N;
    double *w = (double *)malloc(sizeof(double) * n);
    double *Fw = (double *)malloc(sizeof(double) * n);
    double *h = (double *)malloc(sizeof(double) * n);
    if (!w || !Fw || !h) { free(w); free(Fw); free(h); return -2; }
    for (long s = 0; s < n; ++s) h[s] = fd_step(X[s]);

    for (int colour = 0; colour < 21; ++colour) {
        const int cell_colour = colour / 3, dof = colour % 3;
        memcpy(w, X, sizeof(double) * n);
        for (long j = 0; j < ny; ++j)
            for (long i = 0; i < nx; ++i)
                if ((i + 3 * j) % 7 == cell_colour) {
                    long s = 3 * (i + nx * j) + dof;
                    w[s] = X[s] + h[s];
                }
        ldc_oracle_residual(w, U_old, V_old, Fw, nx, ny, dt, dx, dy, rho, mi, bc, use_cds,
                            nthreads);
        for (long r = 0; r < n; ++r) {
            long cr = r / 3;
            for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p) {
                long col = indices[p];
                long cc = col / 3;
                if (col % 3 != dof) continue;
                long ic = cc % nx, jc = cc / nx;
                if ((ic + 3 * jc) % 7 != cell_colour) continue;
                if (is_wrap_pair(cr, cc, nx)) { values[p] = 0.0; continue; }
                values[p] = (Fw[r] - F0[r]) * (1.0 / h[col]);
            }
        }
    }
    free(w); free(Fw); free(h);
    return 0;
}

/* brute-force column-by-column FD (one residual per column) -- used by the tests on
 * tiny grids to prove that the coloured schedule above gives the same values. */
LDC_API int ldc_oracle_fd_jacobian_bycolumn(const double *X, const double *U_old,
                                            const double *V_old, const double *F0,
                                            const int32_t *indptr, const int32_t *indices,
                                            double *values, long nx, long ny, double dt,
                                            double dx, double dy, double rho, double mi,
                                            double bc, int use_cds)
{
    const long N = nx * ny, n = 3 * N;
    double *w = (double *)malloc(sizeof(double) * n);
    double *Fw = (double *)malloc(sizeof(double) * n * 1);
    double *dense_col = (double *)malloc(sizeof(double) * n);
    if (!w || !Fw || !dense_col) { free(w); free(Fw); free(dense_col); return -2; }
    memcpy(w, X, sizeof(double) * n);
    /* column loop outermost: store column results through a row scan */
    for (long col = 0; col < n; ++col) {
        double h = fd_step(X[col]);
        w[col] = X[col] + h;
        ldc_oracle_residual(w, U_old, V_old, Fw, nx, ny, dt, dx, dy, rho, mi, bc, use_cds, 1);
        w[col] = X[col];
        for (long r = 0; r < n; ++r) dense_col[r] = (Fw[r] - F0[r]) * (1.0 / h);
        /* rows that hold this column are within +-(nx+1) cells */
        long cc = col / 3;
        long c_lo = cc - nx - 1 < 0 ? 0 : cc - nx - 1;
        long c_hi = cc + nx + 1 >= N ? N - 1 : cc + nx + 1;
        for (long r = 3 * c_lo; r < 3 * (c_hi + 1); ++r)
            for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p)
                if (indices[p] == col) values[p] = dense_col[r];
    }
    free(w); free(Fw); free(dense_col);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* 5. CSR kernels for the linear solve.  PARITY UNPINNED (PETSc internals).   */
/* ------------------------------------------------------------------------- */
LDC_API void ldc_oracle_spmv(const int32_t *indptr, const int32_t *indices, const double *values,
                             const double *x, double *y, long n, int nthreads)
{
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
    for (long r = 0; r < n; ++r) {
        double s = 0.0;
        for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p) s += values[p] * x[indices[p]];
        y[r] = s;
    }
}

/*
 * ILU(0), natural ordering, on the full mask pattern (explicit zeros take part),
 * PETSc MatLUFactorNumeric_SeqAIJ with pc_factor_shift_type NONZERO
 * (petsc_solver_wrapper.py:141-142): every diagonal gets +shift; when a pivot obeys
 * |pv| <= zeropivot*rowsum the factorisation restarts with shift = shiftamount
 * (first time) or 2*shift.  zeropivot = shiftamount = 100*DBL_EPSILON.
 * Output: lu (same pattern; strictly-lower part = L multipliers, rest = U incl. the
 * diagonal), diag_pos[r] = position of the diagonal in row r.
 * Returns the number of shifts applied (>=0) or <0 on error.
 */
LDC_API int ldc_oracle_ilu0(const int32_t *indptr, const int32_t *indices, const double *values,
                            double *lu, int32_t *diag_pos, long n, double *shift_out)
{
    const double zeropivot = 100.0 * 2.220446049250313e-16;
    const double shiftamount = 100.0 * 2.220446049250313e-16;
    double shift = 0.0;
    int nshift = 0;
    int32_t *where = (int32_t *)malloc(sizeof(int32_t) * n);
    if (!where) return -2;
    for (long r = 0; r < n; ++r) {
        diag_pos[r] = -1;
        for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p)
            if (indices[p] == r) diag_pos[r] = p;
        if (diag_pos[r] < 0) { free(where); return -3; }
    }
    for (long c = 0; c < n; ++c) where[c] = -1;

    for (;;) {
        int newshift = 0;
        for (long r = 0; r < n && !newshift; ++r) {
            const int32_t b = indptr[r], e = indptr[r + 1];
            for (int32_t p = b; p < e; ++p) { lu[p] = values[p]; where[indices[p]] = p; }
            lu[diag_pos[r]] += shift;
            for (int32_t p = b; p < diag_pos[r]; ++p) {
                const long k = indices[p];
                double mult = lu[p];
                if (mult != 0.0) {
                    mult = mult / lu[diag_pos[k]];
                    lu[p] = mult;
                    for (int32_t q = diag_pos[k] + 1; q < indptr[k + 1]; ++q) {
                        int32_t t = where[indices[q]];
                        if (t >= 0) lu[t] -= mult * lu[q];
                    }
                }
            }
            double rs = 0.0;
            for (int32_t p = b; p < e; ++p)
                if (p != diag_pos[r]) rs += fabs(lu[p]);
            const double pv = lu[diag_pos[r]];
            if (fabs(pv) <= zeropivot * rs && pv == pv) {
                shift = (nshift == 0) ? shiftamount : 2.0 * shift;
                ++nshift;
                newshift = 1;
            }
            for (int32_t p = b; p < e; ++p) where[indices[p]] = -1;
        }
        if (!newshift) break;
        if (nshift > 200) { free(where); return -4; }
    }
    free(where);
    if (shift_out) *shift_out = shift;
    return nshift;
}

/* z = (LU)^{-1} r : forward (unit lower) then backward substitution */
LDC_API void ldc_oracle_ilu0_solve(const int32_t *indptr, const int32_t *indices,
                                   const double *lu, const int32_t *diag_pos, const double *r,
                                   double *z, long n)
{
    for (long i = 0; i < n; ++i) {
        double s = r[i];
        for (int32_t p = indptr[i]; p < diag_pos[i]; ++p) s -= lu[p] * z[indices[p]];
        z[i] = s;
    }
    for (long i = n - 1; i >= 0; --i) {
        double s = z[i];
        for (int32_t p = diag_pos[i] + 1; p < indptr[i + 1]; ++p) s -= lu[p] * z[indices[p]];
        z[i] = s / lu[diag_pos[i]];
    }
}

/*
 * GMRES(restart) with LEFT preconditioning M = ILU(0), zero initial guess, classical
 * Gram-Schmidt without refinement, convergence on the preconditioned residual:
 * ||M^-1 r|| <= max(rtol*||M^-1 b||, atol); divergence when it exceeds dtol*||M^-1 b||
 * (PETSc KSP defaults).  Returns the KSP-style reason:
 *   2 rtol, 3 atol, -3 max_it, -4 dtol, -9 NaN.  *iters = total Krylov iterations.
 */
LDC_API int ldc_oracle_gmres_ilu0(const int32_t *indptr, const int32_t *indices,
                                  const double *values, const double *lu,
                                  const int32_t *diag_pos, const double *b, double *x, long n,
                                  int restart, double rtol, double atol, double dtol,
                                  long max_it, long *iters, double *rnorm_out, int nthreads)
{
    const int m = restart;
    double *V = (double *)malloc(sizeof(double) * n * (size_t)(m + 1));
    double *w = (double *)malloc(sizeof(double) * n);
    double *t = (double *)malloc(sizeof(double) * n);
    double *H = (double *)calloc((size_t)(m + 1) * m, sizeof(double));
    double *cs = (double *)calloc(m, sizeof(double));
    double *sn = (double *)calloc(m, sizeof(double));
    double *g = (double *)calloc(m + 1, sizeof(double));
    double *y = (double *)calloc(m, sizeof(double));
    int reason = 0;
    long it = 0;
    double rnorm0 = -1.0, rnorm = 0.0;
    if (!V || !w || !t || !H || !cs || !sn || !g || !y) { reason = -100; goto done; }
    memset(x, 0, sizeof(double) * n);

    while (reason == 0) {
        /* r = M^-1 (b - A x) */
        ldc_oracle_spmv(indptr, indices, values, x, t, n, nthreads);
        for (long i = 0; i < n; ++i) t[i] = b[i] - t[i];
        ldc_oracle_ilu0_solve(indptr, indices, lu, diag_pos, t, V, n);
        double beta = 0.0;
        for (long i = 0; i < n; ++i) beta += V[i] * V[i];
        beta = sqrt(beta);
        if (beta != beta) { reason = -9; break; }
        if (rnorm0 < 0.0) {
            rnorm0 = beta;
            rnorm = beta;
            if (beta <= atol) { reason = 3; break; }
            if (beta == 0.0) { reason = 3; break; }
        }
        for (long i = 0; i < n; ++i) V[i] /= beta;
        memset(g, 0, sizeof(double) * (m + 1));
        g[0] = beta;
        int k = 0;
        for (; k < m && reason == 0; ++k) {
            double *vk = V + (size_t)k * n, *vk1 = V + (size_t)(k + 1) * n;
            ldc_oracle_spmv(indptr, indices, values, vk, t, n, nthreads);
            ldc_oracle_ilu0_solve(indptr, indices, lu, diag_pos, t, w, n);
            /* classical Gram-Schmidt: all dots against the same w, then one update */
            for (int i = 0; i <= k; ++i) {
                const double *vi = V + (size_t)i * n;
                double s = 0.0;
                for (long q = 0; q < n; ++q) s += w[q] * vi[q];
                H[(size_t)i * m + k] = s;
            }
            for (int i = 0; i <= k; ++i) {
                const double *vi = V + (size_t)i * n;
                const double hik = H[(size_t)i * m + k];
                for (long q = 0; q < n; ++q) w[q] -= hik * vi[q];
            }
            double hk1 = 0.0;
            for (long q = 0; q < n; ++q) hk1 += w[q] * w[q];
            hk1 = sqrt(hk1);
            H[(size_t)(k + 1) * m + k] = hk1;
            if (hk1 != 0.0)
                for (long q = 0; q < n; ++q) vk1[q] = w[q] / hk1;
            /* apply stored rotations, then a new one */
            for (int i = 0; i < k; ++i) {
                double a = H[(size_t)i * m + k], c = H[(size_t)(i + 1) * m + k];
                H[(size_t)i * m + k] = cs[i] * a + sn[i] * c;
                H[(size_t)(i + 1) * m + k] = -sn[i] * a + cs[i] * c;
            }
            double a = H[(size_t)k * m + k], c = H[(size_t)(k + 1) * m + k];
            double d = sqrt(a * a + c * c);
            if (d == 0.0) { cs[k] = 1.0; sn[k] = 0.0; }
            else { cs[k] = a / d; sn[k] = c / d; }
            H[(size_t)k * m + k] = cs[k] * a + sn[k] * c;
            H[(size_t)(k + 1) * m + k] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            rnorm = fabs(g[k + 1]);
            ++it;
            if (rnorm != rnorm) reason = -9;
            else if (rnorm <= atol) reason = 3;
            else if (rnorm <= rtol * rnorm0) reason = 2;
            else if (rnorm >= dtol * rnorm0) reason = -4;
            else if (it >= max_it) reason = -3;
            else if (hk1 == 0.0) reason = 2; /* happy breakdown */
        }
        /* x += V_k y, H y = g */
        for (int i = k - 1; i >= 0; --i) {
            double s = g[i];
            for (int j = i + 1; j < k; ++j) s -= H[(size_t)i * m + j] * y[j];
            y[i] = s / H[(size_t)i * m + i];
        }
        for (int i = 0; i < k; ++i) {
            const double *vi = V + (size_t)i * n;
            for (long q = 0; q < n; ++q) x[q] += y[i] * vi[q];
        }
    }
done:
    if (iters) *iters = it;
    if (rnorm_out) *rnorm_out = rnorm;
    free(V); free(w); free(t); free(H); free(cs); free(sn); free(g); free(y);
    return reason;
}
