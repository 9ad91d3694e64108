// Preconditioner of the Krylov solve (sm_100a): additive cell-block relaxation and the
// transfer operators of a geometric multigrid hierarchy on the staggered grid.
//
// Role in the reference: PETSc's PC (pc_type ilu with pc_factor_shift_type NONZERO,
// petsc_solver_wrapper.py:141-143).  ILU(0) in natural ordering is a sequential wavefront
// algorithm and its Krylov iteration count grows ~3x per grid doubling on this saddle-point
// system (oracle probe: 30/100/300 at 32^2/64^2/128^2).  The B200 design replaces it by
//   * a cell-block ("Vanka") relaxation: every pressure cell solves the 5x5 system that
//     couples its pressure with its (up to) four face velocities, momentum rows reduced to
//     their diagonal; all cells are relaxed at once from the same residual (additive /
//     block-Jacobi form, two cells share each face -> weight 1/2), so a sweep is one SpMV plus
//     two streaming passes, fully parallel and deterministic;
//   * a multigrid V-cycle over grids nx/2^l x ny/2^l whose operators are the SAME coloured-FD
//     Jacobian kernel evaluated on the restricted state (finite-volume residuals are integrals,
//     so summing fine residuals and interpolating corrections needs no rescaling).
// Only the converged state has to agree with the reference; iteration counts differ by design.
#include "solver_internal.cuh"

namespace ldc {

// ---------------------------------------------------------------------------------------------
// state restriction: P = mean of the 4 children, U/V = mean of the two fine faces lying on the
// coarse face; coarse dummy slots = 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
restrict_state_kernel(int nxf, int nyf, int v_rows_c, const double *__restrict__ Xf, double *__restrict__ Xc)
{
    const int nxc = nxf / 2, nyc = nyf / 2;
    const long Nc = (long)nxc * nyc;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= Nc) return;
    const int b = blockIdx.y;
    Xf += (long)b * 3 * nxf * nyf;
    Xc += (long)b * 3 * Nc;
    const int J = (int)(c / nxc), I = (int)(c - (long)J * nxc);
    auto f = [&](int i, int j, int d) { return Xf[3 * ((long)i + (long)nxf * j) + d]; };
    const int i0 = 2 * I, j0 = 2 * J;
    Xc[3 * c] = 0.25 * (f(i0, j0, 0) + f(i0 + 1, j0, 0) + f(i0, j0 + 1, 0) + f(i0 + 1, j0 + 1, 0));
    Xc[3 * c + 1] = (I < nxc - 1) ? 0.5 * (f(i0 + 1, j0, 1) + f(i0 + 1, j0 + 1, 1)) : 0.0;
    Xc[3 * c + 2] = (J < v_rows_c) ? 0.5 * (f(i0, j0 + 1, 2) + f(i0 + 1, j0 + 1, 2)) : 0.0;
}

// inverse diagonal of J from the compact layout: U rows -> k=9 (U_P), V rows -> k=20 (V_P);
// continuity rows have no diagonal (0)
template <typename T>
__global__ void __launch_bounds__(256)
extract_diag_kernel(long N, const double *__restrict__ Jc, T *__restrict__ diag)
{
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    const int b = blockIdx.y;
    Jc += (long)b * 26 * N;
    // stored INVERTED: the relaxation only ever divides by the momentum diagonals
    T *d = diag + (long)b * 3 * N + 3 * c;
    const double du = Jc[9 * N + c], dv = Jc[20 * N + c];
    d[0] = (T)0;
    d[1] = (T)((du != 0.0) ? 1.0 / du : 0.0);
    d[2] = (T)((dv != 0.0) ? 1.0 / dv : 0.0);
}

// precision conversion of a flat array (the fp32 copy of the Jacobian for the preconditioner)
template <typename A, typename B>
__global__ void __launch_bounds__(256) convert_kernel(const A *__restrict__ in, B *__restrict__ out, long n)
{
    const long q0 = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (q0 + e < n) out[q0 + e] = (B)in[q0 + e];
}

// ---------------------------------------------------------------------------------------------
// cell-block relaxation.  Local system of cell c=(i,j) with the analytic pressure/divergence
// couplings of the discretisation (residual_function.cpp:77-84,151: dC/dU_e=+dy, dC/dU_w=-dy,
// dC/dV_n=+dx, dC/dV_s=-dx; dFu(c)/dP(c)=-dy, dFu(c-1)/dP(c)=+dy, dFv(c)/dP(c)=-dx,
// dFv(c-nx)/dP(c)=+dx) and the momentum diagonals a_* taken from the FD Jacobian:
//     a_k du_k + g_k dp = r_k  (k = w,e,s,n)      sum_k d_k du_k = r_c
//  => dp = (sum_k d_k r_k / a_k - r_c) / (sum_k d_k g_k / a_k),   du_k = (r_k - g_k dp) / a_k
// ---------------------------------------------------------------------------------------------
// pressure correction of the cell block of cell (i,j); idiag holds 1/a (see extract_diag_kernel)
template <typename T>
__device__ __forceinline__ T vanka_dp(const T *__restrict__ r, const T *__restrict__ idiag,
                                      long c, int i, int j, int nx, int v_rows, T dx, T dy)
{
    // Across a strip boundary the relaxation is decoupled: the south face of the first held row
    // belongs to the strip below and is left out.
    T num = -r[3 * c];
    T den = (T)0;
    if (i < nx - 1) {  // east face: U row of cell c, d=+dy, g=-dy
        const T ia = idiag[3 * c + 1];
        num += dy * r[3 * c + 1] * ia;
        den -= dy * dy * ia;
    }
    if (i > 0) {       // west face: U row of cell c-1, d=-dy, g=+dy
        const T ia = idiag[3 * (c - 1) + 1];
        num -= dy * r[3 * (c - 1) + 1] * ia;
        den -= dy * dy * ia;
    }
    if (j < v_rows) {  // north face: V row of cell c, d=+dx, g=-dx
        const T ia = idiag[3 * c + 2];
        num += dx * r[3 * c + 2] * ia;
        den -= dx * dx * ia;
    }
    if (j > 0) {       // south face: V row of cell c-nx, d=-dx, g=+dx
        const T ia = idiag[3 * (c - nx) + 2];
        num -= dx * r[3 * (c - nx) + 2] * ia;
        den -= dx * dx * ia;
    }
    return (den != (T)0) ? num / den : (T)0;
}

// Two-kernel form for large levels (bandwidth-bound: each block's dp is computed once and stored)
template <typename T>
__global__ void __launch_bounds__(256)
vanka_dp_kernel(int nx, int ny, int v_rows, T dx, T dy, const T *__restrict__ r,
                const T *__restrict__ idiag, T *__restrict__ dp,
                const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const long N = (long)nx * ny;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    const int j = (int)(c / nx), i = (int)(c - (long)j * nx);
    dp[(long)b * N + c] = vanka_dp(r + (long)b * 3 * N, idiag + (long)b * 3 * N, c, i, j, nx, v_rows, dx, dy);
}

template <typename T, bool ASSIGN>
__global__ void __launch_bounds__(256)
vanka_update_kernel(int nx, int ny, int v_rows, T dx, T dy, T om_u, T om_p,
                    const T *__restrict__ r, const T *__restrict__ idiag,
                    const T *__restrict__ dp, T *__restrict__ x,
                    const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const long N = (long)nx * ny;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    r += (long)b * 3 * N;
    idiag += (long)b * 3 * N;
    dp += (long)b * N;
    x += (long)b * 3 * N;
    const int j = (int)(c / nx), i = (int)(c - (long)j * nx);
    const T p_c = dp[c];
    T e0 = om_p * p_c, e1, e2;
    if (i < nx - 1) {
        const T ru = r[3 * c + 1];
        e1 = (T)0.5 * om_u * ((ru + dy * p_c) + (ru - dy * dp[c + 1])) * idiag[3 * c + 1];
    } else {
        e1 = r[3 * c + 1] * idiag[3 * c + 1];
    }
    if (j < v_rows) {
        const T rv = r[3 * c + 2];
        const T p_n = (j + 1 < ny) ? dp[c + nx] : (T)0;
        e2 = (T)0.5 * om_u * ((rv + dx * p_c) + (rv - dx * p_n)) * idiag[3 * c + 2];
    } else {
        e2 = r[3 * c + 2] * idiag[3 * c + 2];
    }
    if (ASSIGN) {
        x[3 * c] = e0;
        x[3 * c + 1] = e1;
        x[3 * c + 2] = e2;
    } else {
        x[3 * c] += e0;
        x[3 * c + 1] += e1;
        x[3 * c + 2] += e2;
    }
}

// Fused form for small (launch-bound) levels.
// One relaxation sweep on a given residual, one thread per cell: the thread needs the pressure
// corrections of its own cell block and of the blocks to the east and north (the two blocks that
// share its U and V face) and simply evaluates all three -- the residual entries come from L1/L2,
// no intermediate array, one launch.  x (+)= omega * correction ; ASSIGN: x is taken as zero.
// ny = held rows, v_rows = held rows with a real V.
template <typename T, bool ASSIGN>
__global__ void __launch_bounds__(256)
vanka_relax_kernel(int nx, int ny, int v_rows, T dx, T dy, T om_u, T om_p,
                   const T *__restrict__ r, const T *__restrict__ idiag,
                   T *__restrict__ x, const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const long N = (long)nx * ny;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    r += (long)b * 3 * N;
    idiag += (long)b * 3 * N;
    x += (long)b * 3 * N;
    const int j = (int)(c / nx), i = (int)(c - (long)j * nx);
    const T p_c = vanka_dp(r, idiag, c, i, j, nx, v_rows, dx, dy);
    T e0 = om_p * p_c, e1, e2;
    if (i < nx - 1) {
        // face shared by cells c (as east face, g=-dy) and c+1 (as west face, g=+dy)
        const T ru = r[3 * c + 1];
        const T p_e = vanka_dp(r, idiag, c + 1, i + 1, j, nx, v_rows, dx, dy);
        e1 = (T)0.5 * om_u * ((ru + dy * p_c) + (ru - dy * p_e)) * idiag[3 * c + 1];
    } else {
        e1 = r[3 * c + 1] * idiag[3 * c + 1];  // dummy row (identity)
    }
    if (j < v_rows) {
        const T rv = r[3 * c + 2];
        // the cell above is in the next strip for the last held row of a strip: decoupled (dp = 0)
        const T p_n = (j + 1 < ny) ? vanka_dp(r, idiag, c + nx, i, j + 1, nx, v_rows, dx, dy) : (T)0;
        e2 = (T)0.5 * om_u * ((rv + dx * p_c) + (rv - dx * p_n)) * idiag[3 * c + 2];
    } else {
        e2 = r[3 * c + 2] * idiag[3 * c + 2];
    }
    if (ASSIGN) {
        x[3 * c] = e0;
        x[3 * c + 1] = e1;
        x[3 * c + 2] = e2;
    } else {
        x[3 * c] += e0;
        x[3 * c + 1] += e1;
        x[3 * c + 2] += e2;
    }
}

// ---------------------------------------------------------------------------------------------
// residual restriction (sum over the fine control volumes covered by the coarse one):
//   continuity: the 4 children;  U rows: weights [1/2,1,1/2] over fine faces 2I,2I+1,2I+2 and
//   both fine rows;  V rows: transposed.  Coarse dummy rows get 0.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
restrict_residual_kernel(int nxf, int nyf, int v_rows_c, const T *__restrict__ rf,
                         T *__restrict__ bc, const int32_t *__restrict__ run)
{
    // a real coarse V row (J < v_rows_c) reads fine row 2J+2, which for the last coarse row of a
    // strip is the halo row above the fine strip (rf carries it)
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const int nxc = nxf / 2, nyc = nyf / 2;
    const long Nc = (long)nxc * nyc;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= Nc) return;
    rf += (long)b * 3 * nxf * nyf;
    bc += (long)b * 3 * Nc;
    const int J = (int)(c / nxc), I = (int)(c - (long)J * nxc);
    auto f = [&](int i, int j, int d) { return rf[3 * ((long)i + (long)nxf * j) + d]; };
    const int i0 = 2 * I, j0 = 2 * J;
    bc[3 * c] = (f(i0, j0, 0) + f(i0 + 1, j0, 0)) + (f(i0, j0 + 1, 0) + f(i0 + 1, j0 + 1, 0));
    T u = (T)0, v = (T)0;
    if (I < nxc - 1) {
        u = (T)0.5 * (f(i0, j0, 1) + f(i0, j0 + 1, 1)) + (f(i0 + 1, j0, 1) + f(i0 + 1, j0 + 1, 1)) +
            (T)0.5 * (f(i0 + 2, j0, 1) + f(i0 + 2, j0 + 1, 1));
    }
    if (J < v_rows_c) {
        v = (T)0.5 * (f(i0, j0, 2) + f(i0 + 1, j0, 2)) + (f(i0, j0 + 1, 2) + f(i0 + 1, j0 + 1, 2)) +
            (T)0.5 * (f(i0, j0 + 2, 2) + f(i0 + 1, j0 + 2, 2));
    }
    bc[3 * c + 1] = u;
    bc[3 * c + 2] = v;
}

// prolongation = transpose of the restriction above, added to the fine correction
template <typename T>
__global__ void __launch_bounds__(256)
prolong_add_kernel(int nxf, int nyf, int v_rows_f, int v_rows_c, int ghost_lo_c,
                   const T *__restrict__ xc, T *__restrict__ xf,
                   const int32_t *__restrict__ run)
{
    // xc carries the halo row below a strip (coarse row -1) when ghost_lo_c is set
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const int nxc = nxf / 2, nyc = nyf / 2;
    const long Nf = (long)nxf * nyf;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= Nf) return;
    xc += (long)b * 3 * nxc * nyc;
    xf += (long)b * 3 * Nf;
    const int j = (int)(c / nxf), i = (int)(c - (long)j * nxf);
    const int I = i >> 1, J = j >> 1;
    auto g = [&](int II, int JJ, int d) { return xc[3 * ((long)II + (long)nxc * JJ) + d]; };
    xf[3 * c] += g(I, J, 0);
    if (i < nxf - 1) {
        T u;
        if (i & 1) u = g(I, J, 1);  // fine face on the coarse face (I <= nxc-2 here)
        else u = (T)0.5 * ((I >= 1 ? g(I - 1, J, 1) : (T)0) + (I < nxc - 1 ? g(I, J, 1) : (T)0));
        xf[3 * c + 1] += u;
    }
    if (j < v_rows_f) {
        T v;
        if (j & 1) v = g(I, J, 2);
        else v = (T)0.5 * (((J >= 1 || ghost_lo_c) ? g(I, J - 1, 2) : (T)0) + (J < v_rows_c ? g(I, J, 2) : (T)0));
        xf[3 * c + 2] += v;
    }
}

// x_P -= mean(x_P): removes the constant-pressure mode (the Jacobian's null space)
// from a coarse-grid correction.  One block per instance; meant for the small coarsest grid.
template <typename T>
__global__ void __launch_bounds__(256)
remove_mean_pressure_kernel(long N, T *__restrict__ x, const int32_t *__restrict__ run)
{
    const int b = blockIdx.x;
    if (run && !run[b]) return;
    x += (long)b * 3 * N;
    __shared__ double sm[256];
    double s = 0.0;
    for (long c = threadIdx.x; c < N; c += 256) s += x[3 * c];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    const T mean = (T)(sm[0] / (double)N);
    for (long c = threadIdx.x; c < N; c += 256) x[3 * c] -= mean;
}

// large grids: chunked partial sums of the pressure slots, fixed-order reduction, subtraction
template <typename T>
__global__ void __launch_bounds__(256)
pressure_sum_partials_kernel(long N, const T *__restrict__ x, double *__restrict__ partials,
                             const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    x += (long)b * 3 * N;
    const long c0 = (long)blockIdx.x * 2048;
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const long c = c0 + threadIdx.x + (long)e * 256;
        if (c < N) s += x[3 * c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sm[k];
        partials[(long)b * gridDim.x + blockIdx.x] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
subtract_mean_pressure_kernel(long N, T *__restrict__ x, const double *__restrict__ sums,
                              double n_cells, const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    if (c < N) x[(long)b * 3 * N + 3 * c] -= (T)(sums[b] / n_cells);
}

static dim3 cell_grid(long N, int batch) { return dim3((unsigned)((N + 255) / 256), batch); }

// two-step form (a distributed solver all-reduces `sums` in between): sums[b] = sum of x_P,
// then x_P -= sums[b] / n_cells_global
template <typename T>
int mg_pressure_sum(const MgLevel &l, const T *x, int batch, const int32_t *run, double *partials,
                    double *sums, cudaStream_t s)
{
    const int nblk = (int)((l.N + 2047) / 2048);
    pressure_sum_partials_kernel<T><<<dim3(nblk, batch), 256, 0, s>>>(l.N, x, partials, run);
    LDC_CHECK_LAUNCH();
    return vec_reduce_partials(partials, 1, nblk, batch, sums, s);
}

template <typename T>
int mg_subtract_mean(const MgLevel &l, T *x, int batch, const int32_t *run, const double *sums,
                     long n_cells_global, cudaStream_t s)
{
    subtract_mean_pressure_kernel<T><<<cell_grid(l.N, batch), 256, 0, s>>>(l.N, x, sums, (double)n_cells_global, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

template <typename T>
int mg_remove_mean_pressure(const MgLevel &l, T *x, int batch, const int32_t *run, double *partials,
                            double *sums, cudaStream_t s)
{
    if (l.N <= 4096 || partials == nullptr) {
        remove_mean_pressure_kernel<T><<<batch, 256, 0, s>>>(l.N, x, run);
        LDC_CHECK_LAUNCH();
        return LDC_OK;
    }
    if (int rc = mg_pressure_sum<T>(l, x, batch, run, partials, sums, s)) return rc;
    return mg_subtract_mean<T>(l, x, batch, run, sums, l.N, s);
}

int mg_restrict_state(const MgLevel &f, const MgLevel &c, int batch, cudaStream_t s)
{
    restrict_state_kernel<<<cell_grid(c.N, batch), 256, 0, s>>>(f.nx, f.ny, c.v_rows, f.X, c.X);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

// inverse diagonals (fp64 and, when the level carries an fp32 copy, fp32 + the fp32 Jacobian)
int mg_extract_diag(const MgLevel &l, int batch, cudaStream_t s)
{
    extract_diag_kernel<double><<<cell_grid(l.N, batch), 256, 0, s>>>(l.N, l.J, l.diag);
    LDC_CHECK_LAUNCH();
    if (l.Jf) {
        extract_diag_kernel<float><<<cell_grid(l.N, batch), 256, 0, s>>>(l.N, l.J, l.diagf);
        LDC_CHECK_LAUNCH();
        const long n = 26 * l.N * batch;
        convert_kernel<double, float><<<(unsigned)((n + 1023) / 1024), 256, 0, s>>>(l.J, l.Jf, n);
        LDC_CHECK_LAUNCH();
    }
    return LDC_OK;
}

template <typename A, typename B>
int vec_convert(const A *in, B *out, long n, cudaStream_t s)
{
    convert_kernel<A, B><<<(unsigned)((n + 1023) / 1024), 256, 0, s>>>(in, out, n);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

// one relaxation on a given residual r with the level's inverse diagonals idiag (dp: scratch)
template <typename T>
int mg_vanka_relax(const MgLevel &l, const T *r, const T *idiag, T *dp, T *x, bool assign, double om_u,
                   double om_p, int batch, const int32_t *run, cudaStream_t s)
{
    const T dx = (T)l.dx, dy = (T)l.dy, ou = (T)om_u, op = (T)om_p;
    if (l.N * batch <= 65536) {   // launch-bound: one fused kernel (measured: 100^2 -45%, 256^2 -25%)
        if (assign)
            vanka_relax_kernel<T, true><<<cell_grid(l.N, batch), 256, 0, s>>>(l.nx, l.ny, l.v_rows, dx, dy, ou, op, r, idiag, x, run);
        else
            vanka_relax_kernel<T, false><<<cell_grid(l.N, batch), 256, 0, s>>>(l.nx, l.ny, l.v_rows, dx, dy, ou, op, r, idiag, x, run);
        LDC_CHECK_LAUNCH();
        return LDC_OK;
    }
    vanka_dp_kernel<T><<<cell_grid(l.N, batch), 256, 0, s>>>(l.nx, l.ny, l.v_rows, dx, dy, r, idiag, dp, run);
    LDC_CHECK_LAUNCH();
    if (assign)
        vanka_update_kernel<T, true><<<cell_grid(l.N, batch), 256, 0, s>>>(l.nx, l.ny, l.v_rows, dx, dy, ou, op, r, idiag, dp, x, run);
    else
        vanka_update_kernel<T, false><<<cell_grid(l.N, batch), 256, 0, s>>>(l.nx, l.ny, l.v_rows, dx, dy, ou, op, r, idiag, dp, x, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

template <typename T>
int mg_restrict_residual(const MgLevel &f, const T *r_f, const MgLevel &c, T *b_c, int batch,
                         const int32_t *run, cudaStream_t s)
{
    restrict_residual_kernel<T><<<cell_grid(c.N, batch), 256, 0, s>>>(f.nx, f.ny, c.v_rows, r_f, b_c, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

template <typename T>
int mg_prolong_add(const MgLevel &c, const T *x_c, const MgLevel &f, T *x_f, int batch,
                   const int32_t *run, cudaStream_t s)
{
    prolong_add_kernel<T><<<cell_grid(f.N, batch), 256, 0, s>>>(f.nx, f.ny, f.v_rows, c.v_rows, c.ghost_lo ? 1 : 0, x_c, x_f, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

// explicit instantiations used by solver.cu
#define LDC_INSTANTIATE(T)                                                                               \
    template int mg_pressure_sum<T>(const MgLevel &, const T *, int, const int32_t *, double *, double *, cudaStream_t); \
    template int mg_subtract_mean<T>(const MgLevel &, T *, int, const int32_t *, const double *, long, cudaStream_t);    \
    template int mg_remove_mean_pressure<T>(const MgLevel &, T *, int, const int32_t *, double *, double *, cudaStream_t); \
    template int mg_vanka_relax<T>(const MgLevel &, const T *, const T *, T *, T *, bool, double, double, int,           \
                                   const int32_t *, cudaStream_t);                                                       \
    template int mg_restrict_residual<T>(const MgLevel &, const T *, const MgLevel &, T *, int, const int32_t *, cudaStream_t); \
    template int mg_prolong_add<T>(const MgLevel &, const T *, const MgLevel &, T *, int, const int32_t *, cudaStream_t);
LDC_INSTANTIATE(double)
LDC_INSTANTIATE(float)
#undef LDC_INSTANTIATE
template int vec_convert<double, float>(const double *, float *, long, cudaStream_t);
template int vec_convert<float, double>(const float *, double *, long, cudaStream_t);

}  // namespace ldc
