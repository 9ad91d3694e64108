// Dense vector kernels of the Krylov solver (sm_100a): fused Gram-Schmidt dot products and
// updates, deterministic two-stage reductions, scaled copies.  Replaces PETSc's
// VecMDot / VecMAXPY / VecNorm / VecScale on sequential vectors.
//
// All kernels are streaming (HBM-bound); each block owns a contiguous chunk of kDotChunk
// elements held in registers (8 per thread, 512 threads), reads every basis vector once with
// coalesced 8-byte loads, and reduces with warp shuffles.  Reductions are summed in a fixed
// order (no atomics), so Krylov iteration counts are reproducible run to run.
#include "solver_internal.cuh"

namespace ldc {

static constexpr int kVT = 512;                 // threads per block
static constexpr int kVE = kDotChunk / kVT;     // elements per thread (8)

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kVT)
multi_dot_kernel(const double *__restrict__ V, long v_stride, int count,
                 const double *__restrict__ w, long n, const int32_t *__restrict__ run,
                 double *__restrict__ partials)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    __shared__ double sm[kMaxRestart + 1][kVT / 32];
    const long base = (long)blockIdx.x * kDotChunk + threadIdx.x;
    const double *wb = w + b * n;
    double wr[kVE];
#pragma unroll
    for (int e = 0; e < kVE; ++e) {
        const long q = base + (long)e * kVT;
        wr[e] = q < n ? wb[q] : 0.0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = 0; i < count; ++i) {
        const double *vi = V + i * v_stride + b * n;
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < kVE; ++e) {
            const long q = base + (long)e * kVT;
            if (q < n) acc = fma(wr[e], __ldg(vi + q), acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) sm[i][warp] = acc;
    }
    __syncthreads();
    if ((int)threadIdx.x < count) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kVT / 32; ++k) t += sm[threadIdx.x][k];
        partials[((long)b * count + threadIdx.x) * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kVT)
gs_update_kernel(const double *__restrict__ V, long v_stride, int count,
                 const double *__restrict__ h, int ldh, double *__restrict__ w, long n,
                 const int32_t *__restrict__ run, double *__restrict__ partials)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    __shared__ double sm[kVT / 32];
    __shared__ double hs[kMaxRestart + 1];
    if ((int)threadIdx.x < count) hs[threadIdx.x] = h[(long)b * ldh + threadIdx.x];
    __syncthreads();
    const long base = (long)blockIdx.x * kDotChunk + threadIdx.x;
    double *wb = w + b * n;
    double wr[kVE];
#pragma unroll
    for (int e = 0; e < kVE; ++e) {
        const long q = base + (long)e * kVT;
        wr[e] = q < n ? wb[q] : 0.0;
    }
    for (int i = 0; i < count; ++i) {
        const double *vi = V + i * v_stride + b * n;
        const double hi = hs[i];
#pragma unroll
        for (int e = 0; e < kVE; ++e) {
            const long q = base + (long)e * kVT;
            if (q < n) wr[e] = fma(-hi, __ldg(vi + q), wr[e]);
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < kVE; ++e) {
        const long q = base + (long)e * kVT;
        if (q < n) {
            wb[q] = wr[e];
            acc = fma(wr[e], wr[e], acc);
        }
    }
    acc = warp_sum(acc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kVT / 32; ++k) t += sm[k];
        partials[(long)b * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kVT)
norm2_partials_kernel(const double *__restrict__ x, long n, const int32_t *__restrict__ run,
                      double *__restrict__ partials)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    __shared__ double sm[kVT / 32];
    const long base = (long)blockIdx.x * kDotChunk + threadIdx.x;
    const double *xb = x + b * n;
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < kVE; ++e) {
        const long q = base + (long)e * kVT;
        if (q < n) {
            const double v = xb[q];
            acc = fma(v, v, acc);
        }
    }
    acc = warp_sum(acc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kVT / 32; ++k) t += sm[k];
        partials[(long)b * gridDim.x + blockIdx.x] = t;
    }
}

// grid (count, batch): one block sums nblk partials in a fixed order
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double *__restrict__ partials, int count, int nblk,
                       double *__restrict__ out)
{
    const int i = blockIdx.x, b = blockIdx.y;
    const double *q = partials + ((long)b * count + i) * nblk;
    double s = 0.0;
    for (int k = threadIdx.x; k < nblk; k += 256) s += q[k];
    __shared__ double sm[256];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(long)b * count + i] = sm[0];
}

__global__ void __launch_bounds__(256)
scale_dev_kernel(const double *x, const double *__restrict__ alpha, int lda, bool inv,
                 double *y, long n, const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    double a = alpha[(long)b * lda];
    if (inv) a = (a != 0.0) ? 1.0 / a : 0.0;
    const long q0 = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
    const double *xb = x + b * n;
    double *yb = y + b * n;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (q0 + e < n) yb[q0 + e] = a * xb[q0 + e];
}

__global__ void __launch_bounds__(kVT)
multi_axpy_kernel(const double *__restrict__ Z, long z_stride, int count,
                  const double *__restrict__ y, int ldy, double *__restrict__ x, long n,
                  const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    __shared__ double ys[kMaxRestart + 1];
    if ((int)threadIdx.x < count) ys[threadIdx.x] = y[(long)b * ldy + threadIdx.x];
    __syncthreads();
    const long base = (long)blockIdx.x * kDotChunk + threadIdx.x;
    double *xb = x + b * n;
    double xr[kVE];
#pragma unroll
    for (int e = 0; e < kVE; ++e) {
        const long q = base + (long)e * kVT;
        xr[e] = q < n ? xb[q] : 0.0;
    }
    for (int i = 0; i < count; ++i) {
        const double *zi = Z + i * z_stride + b * n;
        const double yi = ys[i];
#pragma unroll
        for (int e = 0; e < kVE; ++e) {
            const long q = base + (long)e * kVT;
            if (q < n) xr[e] = fma(yi, __ldg(zi + q), xr[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < kVE; ++e) {
        const long q = base + (long)e * kVT;
        if (q < n) xb[q] = xr[e];
    }
}

__global__ void __launch_bounds__(256)
axpby_kernel(double a, const double *x, double bb, const double *y,
             double *z, long n, const int32_t *__restrict__ run)
{
    const int b = blockIdx.y;
    if (run && !run[b]) return;
    const long q0 = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long q = q0 + e;
        if (q < n) {
            double v = 0.0;
            if (x) v = a * x[b * n + q];
            if (y) v = fma(bb, y[b * n + q], v);
            z[b * n + q] = v;
        }
    }
}

// max |U - U_old| over real U slots; grid (blocks, batch) with atomicMax on the bit pattern
// (non-negative doubles order like unsigned integers; max is order independent -> deterministic)
__global__ void __launch_bounds__(256)
du_inf_kernel(const double *__restrict__ x, const double *__restrict__ U_old, int nx, int ny,
              unsigned long long *__restrict__ out)
{
    const int b = blockIdx.y;
    const long nU = (long)(nx - 1) * ny;
    const long k = (long)blockIdx.x * 256 + threadIdx.x;
    double v = 0.0;
    if (k < nU) {
        const long j = k / (nx - 1), i = k - j * (nx - 1);
        v = fabs(x[(long)b * 3 * nx * ny + 3 * (i + (long)nx * j) + 1] - U_old[b * nU + k]);
        if (v != v) v = __longlong_as_double(0x7ff0000000000000LL);  // NaN -> +inf
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out + b, (unsigned long long)__double_as_longlong(v));
}

static dim3 chunk_grid(long n, int batch) { return dim3((unsigned)dot_blocks(n), batch); }
static dim3 flat_grid(long n, int batch) { return dim3((unsigned)((n + 1023) / 1024), batch); }

int vec_multi_dot(const double *V, long v_stride, int count, const double *w, long n, int batch,
                  const int32_t *run, double *partials, cudaStream_t s)
{
    multi_dot_kernel<<<chunk_grid(n, batch), kVT, 0, s>>>(V, v_stride, count, w, n, run, partials);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

int vec_gs_update(const double *V, long v_stride, int count, const double *h, int ldh, double *w,
                  long n, int batch, const int32_t *run, double *partials, cudaStream_t s)
{
    gs_update_kernel<<<chunk_grid(n, batch), kVT, 0, s>>>(V, v_stride, count, h, ldh, w, n, run, partials);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

int vec_reduce_partials(const double *partials, int count, int nblk, int batch, double *out,
                        cudaStream_t s)
{
    reduce_partials_kernel<<<dim3(count, batch), 256, 0, s>>>(partials, count, nblk, out);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

int vec_scale_dev(const double *x, const double *alpha, int lda, bool inv, double *y, long n,
                  int batch, const int32_t *run, cudaStream_t s)
{
    scale_dev_kernel<<<flat_grid(n, batch), 256, 0, s>>>(x, alpha, lda, inv, y, n, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

int vec_multi_axpy(const double *Z, long z_stride, int count, const double *y, int ldy, double *x,
                   long n, int batch, const int32_t *run, cudaStream_t s)
{
    multi_axpy_kernel<<<chunk_grid(n, batch), kVT, 0, s>>>(Z, z_stride, count, y, ldy, x, n, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

int vec_axpby(double a, const double *x, double b, const double *y, double *z, long n, int batch,
              const int32_t *run, cudaStream_t s)
{
    axpby_kernel<<<flat_grid(n, batch), 256, 0, s>>>(a, x, b, y, z, n, run);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

int vec_copy(const double *x, double *y, long n, int batch, const int32_t *run, cudaStream_t s)
{
    return vec_axpby(1.0, x, 0.0, nullptr, y, n, batch, run, s);
}

int vec_zero(double *x, long n, int batch, const int32_t *run, cudaStream_t s)
{
    return vec_axpby(0.0, nullptr, 0.0, nullptr, x, n, batch, run, s);
}

int vec_norm2_partials(const double *x, long n, int batch, const int32_t *run, double *partials,
                       cudaStream_t s)
{
    norm2_partials_kernel<<<chunk_grid(n, batch), kVT, 0, s>>>(x, n, run, partials);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

int vec_du_inf(const double *x, const double *U_old, int nx, int ny, int batch, double *out,
               cudaStream_t s)
{
    LDC_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * batch, s));
    const long nU = (long)(nx - 1) * ny;
    du_inf_kernel<<<dim3((unsigned)((nU + 255) / 256), batch), 256, 0, s>>>(
        x, U_old, nx, ny, (unsigned long long *)out);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

}  // namespace ldc
