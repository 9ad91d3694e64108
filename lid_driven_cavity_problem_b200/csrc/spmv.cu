// Sparse matrix-vector products on the Jacobian mask (sm_100a).
//
// Replaces PETSc MatMult_SeqAIJ on the matrix the reference hands over through
// PETSc.Mat().createAIJWithArrays (nonlinear_solver/petsc_solver_wrapper.py:116-117).
//
//   spmv_mask_kernel : the matrix is the CSR data array of the mask, but the structure is
//                      closed-form (MaskGeom), so NO index array is read: 63*8 B of values per
//                      cell stream in fully coalesced, the x entries of the three cell-row
//                      segments a block touches are staged once in shared memory, and every
//                      row sums its 21 products in CSR column order (the same summation order
//                      as a textbook CSR loop, so results match the CPU product bit for bit
//                      when that one does not fuse multiply-add).
//   spmv_csr_kernel  : canonical CSR (indptr/indices/values) for callers that own an explicit
//                      pattern; same flat coalesced scheme with indices read from memory.
#include "solver_internal.cuh"

namespace ldc {

static constexpr int kSpmvCells = 32;             // cells per block
static constexpr int kSpmvThreads = 128;
static constexpr int kSpmvVals = kSpmvCells * 63;  // 2016 values per interior block

// shared x layout: three segments of consecutive cells (3 doubles each)
//   lo : cells [c0-nx,   c0-nx+32]   (offsets -nx, -nx+1)      33 cells
//   mid: cells [c0-1,    c0+32]      (offsets -1, 0, +1)       34 cells
//   hi : cells [c0+nx-1, c0+nx+31]   (offsets nx-1, nx)        33 cells
static constexpr int kSegLo = 0, kSegMid = 99, kSegHi = 201, kXs = 300;

__device__ __forceinline__ int xs_base(int k)
{   // segment base + 3*shift of mask offset k (ascending offsets)
    switch (k) {
        case 0: return kSegLo;
        case 1: return kSegLo + 3;
        case 2: return kSegMid;
        case 3: return kSegMid + 3;
        case 4: return kSegMid + 6;
        case 5: return kSegHi;
        default: return kSegHi + 3;
    }
}

__global__ void __launch_bounds__(kSpmvThreads)
spmv_mask_kernel(int nx, int ny, long nnz, const double *__restrict__ values,
                 const double *__restrict__ x, const double *__restrict__ rhs,
                 double *__restrict__ y, const int32_t *__restrict__ run)
{
    if (run && !run[blockIdx.y]) return;
    __shared__ double prod[kSpmvVals];
    __shared__ double xs[kXs];

    const long N = (long)nx * ny;
    const int inst = blockIdx.y;
    values += inst * nnz;
    x += inst * 3 * N;
    y += inst * 3 * N;
    if (rhs) rhs += inst * 3 * N;
    const MaskGeom m{nx, N};
    const long c0 = (long)blockIdx.x * kSpmvCells;
    const long c1 = (c0 + kSpmvCells < N) ? c0 + kSpmvCells : N;
    const int tid = threadIdx.x;

    // ---- stage x
    for (int t = tid; t < kXs; t += kSpmvThreads) {
        long cell;
        int e;
        if (t < kSegMid) { cell = c0 - nx + t / 3; e = t % 3; }
        else if (t < kSegHi) { cell = c0 - 1 + (t - kSegMid) / 3; e = (t - kSegMid) % 3; }
        else { cell = c0 + nx - 1 + (t - kSegHi) / 3; e = (t - kSegHi) % 3; }
        xs[t] = (cell >= 0 && cell < N) ? __ldg(x + 3 * cell + e) : 0.0;
    }
    __syncthreads();

    const bool interior = (c0 >= nx) && (c1 == c0 + kSpmvCells) && (c1 - 1 + nx < N);
    if (interior) {
        const double *v = values + 9 * m.pairs_before(c0);  // == 63*c0 - const
        // issue all value loads first (16 independent 8-byte loads per thread in flight)
        double vv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int t = tid + u * kSpmvThreads;
            vv[u] = (t < kSpmvVals) ? __ldg(v + t) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int t = tid + u * kSpmvThreads;
            if (t < kSpmvVals) {
                const int cl = t / 63, rem = t - cl * 63;
                const int ke = rem % 21;
                const int k = ke / 3, e = ke - 3 * k;
                prod[t] = __dmul_rn(vv[u], xs[xs_base(k) + 3 * cl + e]);
            }
        }
        __syncthreads();
        if (tid < 3 * kSpmvCells) {
            const double *q = prod + tid * 21;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 21; ++k) s = __dadd_rn(s, q[k]);
            y[3 * c0 + tid] = rhs ? __dsub_rn(rhs[3 * c0 + tid], s) : s;
        }
    } else {
        // boundary blocks (first/last nx cells): per-row loops over the existing offsets
        const int r = tid;
        const long c = c0 + r / 3;
        if (r < 3 * kSpmvCells && c < c1) {
            const int d = r % 3, cl = r / 3;
            const int lo = m.lo_missing(c), cnt = m.count(c);
            const double *v = values + m.row_start(c, d);
            double s = 0.0;
            for (int kk = 0; kk < cnt; ++kk) {
                const int base = xs_base(lo + kk) + 3 * cl;
#pragma unroll
                for (int e = 0; e < 3; ++e)
                    s = __dadd_rn(s, __dmul_rn(__ldg(v + 3 * kk + e), xs[base + e]));
            }
            y[3 * c + d] = rhs ? __dsub_rn(rhs[3 * c + d], s) : s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// compact layout (26 true entries per cell, SoA Jc[k*N + c], see jacobian.cu): one thread per
// cell, 26 fully coalesced coefficient loads in flight per thread, the three x row segments a
// block of 256 consecutive cells touches staged once in shared memory, y staged so the global
// store is contiguous.  208 + 48 B/cell instead of 504 + 48.  Used by the Krylov solver and
// the multigrid smoother (same numbers as the CSR product: the dropped entries are exact zeros).
// ---------------------------------------------------------------------------------------------
#ifndef LDC_CMP_CELLS
#define LDC_CMP_CELLS 128
#endif
static constexpr int kCmpCells = LDC_CMP_CELLS;
static constexpr int kCmpLo = 0, kCmpMid = 3 * (kCmpCells + 1), kCmpHi = 3 * (2 * kCmpCells + 3),
                     kCmpXs = 3 * (3 * kCmpCells + 4);

template <typename T>
__global__ void __launch_bounds__(kCmpCells)
spmv_compact_kernel(int nx, int ny, const T *__restrict__ Jc, const T *__restrict__ x,
                    const T *__restrict__ rhs, T *__restrict__ y,
                    const int32_t *__restrict__ run, int ghost_lo, int ghost_hi)
{
    // ny = rows of this block of cells (a whole cavity or a row strip).  For a strip, x points at
    // the first owned row and ghost_lo / ghost_hi say whether the row below / above it holds valid
    // halo data; without them the product is the strip's diagonal block (block-Jacobi).
    if (run && !run[blockIdx.y]) return;
    __shared__ T xs[kCmpXs];
    __shared__ T ys[3 * kCmpCells];
    const long N = (long)nx * ny;
    const int inst = blockIdx.y;
    Jc += (long)inst * 26 * N;
    x += (long)inst * 3 * N;
    y += (long)inst * 3 * N;
    if (rhs) rhs += (long)inst * 3 * N;
    const long c0 = (long)blockIdx.x * kCmpCells;
    const int tid = threadIdx.x;
    const long c = c0 + tid;
    const long cell_lo = ghost_lo ? -(long)nx : 0, cell_hi = N + (ghost_hi ? nx : 0);

    // coefficient loads first: they are independent of the staging below
    T J[26];
    if (c < N) {
#pragma unroll
        for (int k = 0; k < 26; ++k) J[k] = __ldg(Jc + (long)k * N + c);
    } else {
#pragma unroll
        for (int k = 0; k < 26; ++k) J[k] = (T)0;
    }
    for (int t = tid; t < kCmpXs; t += kCmpCells) {
        long cell;
        int e;
        if (t < kCmpMid) { cell = c0 - nx + t / 3; e = t % 3; }
        else if (t < kCmpHi) { cell = c0 - 1 + (t - kCmpMid) / 3; e = (t - kCmpMid) % 3; }
        else { cell = c0 + nx - 1 + (t - kCmpHi) / 3; e = (t - kCmpHi) % 3; }
        xs[t] = (cell >= cell_lo && cell < cell_hi) ? __ldg(x + 3 * cell + e) : (T)0;
    }
    __syncthreads();
    {
        const T *lo = xs + kCmpLo + 3 * tid;     // cell c-nx   (lo[3..5] = c-nx+1)
        const T *mi = xs + kCmpMid + 3 * tid;    // cell c-1    (mi[3..5] = c, mi[6..8] = c+1)
        const T *hi = xs + kCmpHi + 3 * tid;     // cell c+nx-1 (hi[3..5] = c+nx)
        T s0 = J[0] * lo[2];
        s0 = fma(J[1], mi[1], s0);
        s0 = fma(J[2], mi[4], s0);
        s0 = fma(J[3], mi[5], s0);
        T s1 = J[4] * lo[1];
        s1 = fma(J[5], lo[2], s1);
        s1 = fma(J[6], lo[5], s1);
        s1 = fma(J[7], mi[1], s1);
        s1 = fma(J[8], mi[3], s1);
        s1 = fma(J[9], mi[4], s1);
        s1 = fma(J[10], mi[5], s1);
        s1 = fma(J[11], mi[6], s1);
        s1 = fma(J[12], mi[7], s1);
        s1 = fma(J[13], mi[8], s1);
        s1 = fma(J[14], hi[4], s1);
        T s2 = J[15] * lo[2];
        s2 = fma(J[16], mi[1], s2);
        s2 = fma(J[17], mi[2], s2);
        s2 = fma(J[18], mi[3], s2);
        s2 = fma(J[19], mi[4], s2);
        s2 = fma(J[20], mi[5], s2);
        s2 = fma(J[21], mi[8], s2);
        s2 = fma(J[22], hi[1], s2);
        s2 = fma(J[23], hi[3], s2);
        s2 = fma(J[24], hi[4], s2);
        s2 = fma(J[25], hi[5], s2);
        ys[3 * tid] = s0;
        ys[3 * tid + 1] = s1;
        ys[3 * tid + 2] = s2;
    }
    __syncthreads();
    const long base = 3 * c0;
    const long limit = 3 * N;
#pragma unroll
    for (int t = tid; t < 3 * kCmpCells; t += kCmpCells) {
        const long q = base + t;
        if (q < limit) y[q] = rhs ? rhs[q] - ys[t] : ys[t];
    }
}

// ---------------------------------------------------------------------------------------------
static constexpr int kCsrRows = 128;
static constexpr int kCsrCap = 128 * 24;

__global__ void __launch_bounds__(kCsrRows)
spmv_csr_kernel(long n_rows, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                const double *__restrict__ values, const double *__restrict__ x,
                double *__restrict__ y)
{
    __shared__ double prod[kCsrCap];
    __shared__ int32_t rp[kCsrRows + 1];
    const long r0 = (long)blockIdx.x * kCsrRows;
    const int nr = (int)((r0 + kCsrRows < n_rows ? r0 + kCsrRows : n_rows) - r0);
    const int tid = threadIdx.x;
    if (tid < nr) rp[tid] = indptr[r0 + tid];
    if (tid == 0) rp[nr] = indptr[r0 + nr];
    __syncthreads();
    const int base = rp[0];
    const int len = rp[nr] - base;
    if (len <= kCsrCap) {
        for (int t = tid; t < len; t += kCsrRows)
            prod[t] = __dmul_rn(__ldg(values + base + t), __ldg(x + __ldg(indices + base + t)));
        __syncthreads();
        if (tid < nr) {
            double s = 0.0;
            for (int q = rp[tid] - base; q < rp[tid + 1] - base; ++q) s = __dadd_rn(s, prod[q]);
            y[r0 + tid] = s;
        }
    } else if (tid < nr) {  // long rows: plain per-row loop
        double s = 0.0;
        for (int q = rp[tid]; q < rp[tid + 1]; ++q)
            s = __dadd_rn(s, __dmul_rn(__ldg(values + q), __ldg(x + __ldg(indices + q))));
        y[r0 + tid] = s;
    }
}

}  // namespace ldc

namespace ldc {
int spmv_mask_launch(int nx, int ny, int batch, const double *values, const double *x,
                     const double *b, double *y, const int32_t *run, cudaStream_t s)
{
    const long N = (long)nx * ny;
    dim3 grid((unsigned)((N + kSpmvCells - 1) / kSpmvCells), batch);
    spmv_mask_kernel<<<grid, kSpmvThreads, 0, s>>>(nx, ny, ldc_mask_nnz(nx, ny, 3), values, x, b, y, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}
template <typename T>
int spmv_compact_launch(int nx, int ny, int batch, const T *Jc, const T *x, const T *b, T *y,
                        const int32_t *run, cudaStream_t s, bool ghost_lo, bool ghost_hi)
{
    const long N = (long)nx * ny;
    dim3 grid((unsigned)((N + kCmpCells - 1) / kCmpCells), batch);
    spmv_compact_kernel<T><<<grid, kCmpCells, 0, s>>>(nx, ny, Jc, x, b, y, run, ghost_lo ? 1 : 0, ghost_hi ? 1 : 0);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}
template int spmv_compact_launch<double>(int, int, int, const double *, const double *, const double *, double *,
                                         const int32_t *, cudaStream_t, bool, bool);
template int spmv_compact_launch<float>(int, int, int, const float *, const float *, const float *, float *,
                                        const int32_t *, cudaStream_t, bool, bool);
}  // namespace ldc

using namespace ldc;

extern "C" int ldc_spmv_compact(const ldc_grid *g, const double *compact_dev, const double *x_dev,
                                double *y_dev, void *stream)
{
    if (int rc = validate_grid(g, false)) return rc;
    LDC_REQUIRE(compact_dev && x_dev && y_dev && x_dev != y_dev, "ldc_spmv_compact: bad device pointers");
    // row strip: x_dev points at the first owned row, halo rows (if the strip has neighbours) around it
    return spmv_compact_launch<double>(g->nx, g->row_count, g->batch, compact_dev, x_dev, nullptr, y_dev, nullptr,
                               (cudaStream_t)stream, g->row_begin > 0, g->row_begin + g->row_count < g->ny);
}

extern "C" int ldc_spmv_mask(const ldc_grid *g, const double *values_dev, const double *x_dev,
                             double *y_dev, void *stream)
{
    if (int rc = validate_grid(g, true)) return rc;
    LDC_REQUIRE(values_dev && x_dev && y_dev && x_dev != y_dev, "ldc_spmv_mask: bad device pointers");
    return spmv_mask_launch(g->nx, g->ny, g->batch, values_dev, x_dev, nullptr, y_dev, nullptr,
                            (cudaStream_t)stream);
}

extern "C" int ldc_spmv_csr(int64_t n_rows, const int32_t *indptr_dev, const int32_t *indices_dev,
                            const double *values_dev, const double *x_dev, double *y_dev, void *stream)
{
    LDC_REQUIRE(n_rows > 0 && indptr_dev && indices_dev && values_dev && x_dev && y_dev && x_dev != y_dev,
                "ldc_spmv_csr: bad arguments");
    spmv_csr_kernel<<<(unsigned)((n_rows + kCsrRows - 1) / kCsrRows), kCsrRows, 0, (cudaStream_t)stream>>>(
        n_rows, indptr_dev, indices_dev, values_dev, x_dev, y_dev);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}
