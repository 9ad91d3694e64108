// Launch wrappers shared by the Newton-Krylov solver (solver.cu), its vector kernels
// (vector_ops.cu) and the multigrid / cell-block preconditioner (precond.cu).
//
// Batch convention: every vector holds `batch` instances back to back (stride n).  `run` is a
// device array of `batch` int32 flags; an instance whose flag is 0 is skipped by every kernel,
// which is how converged / diverged / inactive cavities of a Reynolds ensemble drop out of the
// lock-step iteration without a host round trip.
#pragma once
#include "ldc_internal.cuh"

namespace ldc {

static constexpr int kMaxRestart = 64;   // upper bound of gmres_restart
static constexpr int kDotChunk = 4096;   // vector elements per block in the reduction kernels

inline int dot_blocks(long n) { return (int)((n + kDotChunk - 1) / kDotChunk); }

// ---- vector_ops.cu -------------------------------------------------------------------------
// partials[(b*count + i)*nblk + blk] = sum over the block's chunk of w[b] * V_i[b],  i < count
// V_i = V + i*v_stride (v_stride >= batch*n)
int vec_multi_dot(const double *V, long v_stride, int count, const double *w, long n, int batch,
                  const int32_t *run, double *partials, cudaStream_t s);
// w -= sum_i h[b*ldh + i] V_i ; then partials[b*nblk + blk] = sum w^2 over the chunk
int vec_gs_update(const double *V, long v_stride, int count, const double *h, int ldh, double *w,
                  long n, int batch, const int32_t *run, double *partials, cudaStream_t s);
// out[b*count + i] = sum_blk partials[(b*count+i)*nblk + blk]   (fixed order)
int vec_reduce_partials(const double *partials, int count, int nblk, int batch, double *out,
                        cudaStream_t s);
// y = alpha[b*lda] * x  (alpha read on the device; inv=true uses 1/alpha)
int vec_scale_dev(const double *x, const double *alpha, int lda, bool inv, double *y, long n,
                  int batch, const int32_t *run, cudaStream_t s);
// x += sum_i y[b*ldy + i] * Z_i
int vec_multi_axpy(const double *Z, long z_stride, int count, const double *y, int ldy, double *x,
                   long n, int batch, const int32_t *run, cudaStream_t s);
// z = a*x + b*y   (host scalars)
int vec_axpby(double a, const double *x, double b, const double *y, double *z, long n, int batch,
              const int32_t *run, cudaStream_t s);
int vec_copy(const double *x, double *y, long n, int batch, const int32_t *run, cudaStream_t s);
int vec_zero(double *x, long n, int batch, const int32_t *run, cudaStream_t s);
// partials for sum(x^2) (count = 1 layout), to be reduced by vec_reduce_partials
int vec_norm2_partials(const double *x, long n, int batch, const int32_t *run, double *partials,
                       cudaStream_t s);
// out[b] = max_i |U(x) - U_old| over the real U slots of the interleaved vector x
int vec_du_inf(const double *x, const double *U_old, int nx, int ny, int batch, double *out,
               cudaStream_t s);

// ---- jacobian.cu ---------------------------------------------------------------------------
// coloured-FD Jacobian into the CSR(mask) value array and/or the compact 26-entry SoA layout
int fd_jacobian_launch(const ldc_grid *g, const double *X, const double *U_old, const double *V_old,
                       double *csr_values, double *compact, cudaStream_t st);

// ---- spmv.cu (internal entry with residual form) -------------------------------------------
// y = J x            (b == nullptr)
// y = b - J x        (b != nullptr)
int spmv_mask_launch(int nx, int ny, int batch, const double *values, const double *x,
                     const double *b, double *y, const int32_t *run, cudaStream_t s);
// same on the compact layout Jc[26][N] (see jacobian.cu)
// ny = rows of the block of cells; ghost_lo/ghost_hi: x carries valid halo rows below/above
template <typename T>
int spmv_compact_launch(int nx, int ny, int batch, const T *Jc, const T *x, const T *b, T *y,
                        const int32_t *run, cudaStream_t s, bool ghost_lo = false, bool ghost_hi = false);

// ---- precond.cu ----------------------------------------------------------------------------
struct MgLevel {
    int nx, ny;          // cells per row, rows HELD (a whole cavity, or the rows of this rank's strip)
    // row-strip geometry of the level (a whole cavity: row_begin=0, ny_global=ny, no ghosts)
    int row_begin, ny_global, v_rows;   // v_rows = held rows whose V is a real unknown
    bool ghost_lo, ghost_hi;            // a strip below / above exists (vectors carry halo rows)
    // coarse levels of a distributed hierarchy are REPLICATED: every rank holds the whole (small)
    // level and works on it without communication; slice_* is this rank's share of its rows
    bool replicated;
    int slice_begin, slice_rows;
    double dx, dy;
    long N, n, nnz;
    double *X;       // state the level's Jacobian is linearised about   [batch*n]
    double *J;       // compact Jacobian Jc[26][N] (jacobian.cu)            [batch*26*N]
    double *diag;    // INVERSE diagonal of J (0 for continuity rows)     [batch*n]
    double *x, *b, *r;  // correction, right-hand side, residual          [batch*n]
    double *dp;      // cell pressure corrections of the relaxation       [batch*N]
    // fp32 twins for the mixed-precision cycle (nullptr when the solver runs the cycle in fp64)
    float *Jf, *diagf, *xf, *bf, *rf, *dpf;
};

// Vectors of a level (X, x, b, r) are allocated with one spare row below and above the held rows;
// for a strip those rows receive the neighbours' boundary rows (halo) before an operator reads them.
// The relaxation / transfer wrappers are templates on the scalar type: the multigrid cycle of a
// non-distributed solver runs in fp32 (it is only a preconditioner of the flexible GMRES, which
// stays in fp64), halving the traffic of its bandwidth-bound levels.
int mg_restrict_state(const MgLevel &f, const MgLevel &c, int batch, cudaStream_t s);
int mg_extract_diag(const MgLevel &l, int batch, cudaStream_t s);
template <typename A, typename B> int vec_convert(const A *in, B *out, long n, cudaStream_t s);
template <typename T>
int mg_vanka_relax(const MgLevel &l, const T *r, const T *idiag, T *dp, T *x, bool assign, double om_u,
                   double om_p, int batch, const int32_t *run, cudaStream_t s);
template <typename T>
int mg_restrict_residual(const MgLevel &f, const T *r_f, const MgLevel &c, T *b_c, int batch,
                         const int32_t *run, cudaStream_t s);
template <typename T>
int mg_prolong_add(const MgLevel &c, const T *x_c, const MgLevel &f, T *x_f, int batch,
                   const int32_t *run, cudaStream_t s);
// x_P -= mean(x_P) per instance; partials/sums are scratch (>= batch*ceil(N/2048) and batch doubles)
template <typename T>
int mg_remove_mean_pressure(const MgLevel &l, T *x, int batch, const int32_t *run, double *partials,
                            double *sums, cudaStream_t s);
template <typename T>
int mg_pressure_sum(const MgLevel &l, const T *x, int batch, const int32_t *run, double *partials,
                    double *sums, cudaStream_t s);
template <typename T>
int mg_subtract_mean(const MgLevel &l, T *x, int batch, const int32_t *run, const double *sums,
                     long n_cells_global, cudaStream_t s);

}  // namespace ldc
