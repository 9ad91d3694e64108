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
}  // namespace ldc

using namespace ldc;

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
