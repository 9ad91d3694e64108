// Packing between the three meshes and the interleaved unknown vector (device side of
// _create_X / _recover_X, nonlinear_solver/_common.py:4-36).
#include "ldc_internal.cuh"

namespace ldc {

__global__ void __launch_bounds__(256)
pack_kernel(int nx, int ny, const double *__restrict__ U, const double *__restrict__ V,
            const double *__restrict__ P, double *__restrict__ X)
{
    const long N = (long)nx * ny;
    const long b = blockIdx.y;
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int j = (int)(c / nx), i = (int)(c - (long)j * nx);
    double *x = X + 3 * (b * N + c);
    x[0] = P[b * N + c];
    x[1] = (i < nx - 1) ? U[b * (long)(nx - 1) * ny + i + (long)(nx - 1) * j] : 0.0;
    x[2] = (j < ny - 1) ? V[b * (long)nx * (ny - 1) + c] : 0.0;
}

__global__ void __launch_bounds__(256)
unpack_kernel(int nx, int ny, const double *__restrict__ X, double *__restrict__ U,
              double *__restrict__ V, double *__restrict__ P)
{
    const long N = (long)nx * ny;
    const long b = blockIdx.y;
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int j = (int)(c / nx), i = (int)(c - (long)j * nx);
    const double *x = X + 3 * (b * N + c);
    P[b * N + c] = x[0];
    if (i < nx - 1) U[b * (long)(nx - 1) * ny + i + (long)(nx - 1) * j] = x[1];
    if (j < ny - 1) V[b * (long)nx * (ny - 1) + c] = x[2];
}

}  // namespace ldc

using namespace ldc;

extern "C" int ldc_pack_X(const ldc_grid *g, const double *U_dev, const double *V_dev,
                          const double *P_dev, double *X_dev, void *stream)
{
    if (int rc = validate_grid(g, true)) return rc;
    LDC_REQUIRE(U_dev && V_dev && P_dev && X_dev, "ldc_pack_X: null device pointer");
    const long N = (long)g->nx * g->ny;
    dim3 grid((unsigned)((N + 255) / 256), g->batch);
    pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g->nx, g->ny, U_dev, V_dev, P_dev, X_dev);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}

extern "C" int ldc_unpack_X(const ldc_grid *g, const double *X_dev, double *U_dev, double *V_dev,
                            double *P_dev, void *stream)
{
    if (int rc = validate_grid(g, true)) return rc;
    LDC_REQUIRE(U_dev && V_dev && P_dev && X_dev, "ldc_unpack_X: null device pointer");
    const long N = (long)g->nx * g->ny;
    dim3 grid((unsigned)((N + 255) / 256), g->batch);
    unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g->nx, g->ny, X_dev, U_dev, V_dev, P_dev);
    LDC_CHECK_LAUNCH();
    return LDC_OK;
}
