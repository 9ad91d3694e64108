// Error plumbing, argument validation and device queries of libldc_b200.so.
#include <stdarg.h>
#include <string.h>

#include "ldc_internal.cuh"

namespace ldc {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return LDC_ERR_CUDA;
}

int validate_grid(const ldc_grid *g, bool need_full_grid)
{
    LDC_REQUIRE(g != nullptr, "grid descriptor is NULL");
    LDC_REQUIRE(g->nx >= 3 && g->ny >= 2, "grid %dx%d too small (need nx>=3, ny>=2)", g->nx, g->ny);
    LDC_REQUIRE((long)g->nx * g->ny <= 715827882L, "grid %dx%d: 3*nx*ny exceeds int32", g->nx, g->ny);
    LDC_REQUIRE(g->batch >= 1, "batch must be >= 1 (got %d)", g->batch);
    LDC_REQUIRE(g->row_begin >= 0 && g->row_count >= 1 && g->row_begin + g->row_count <= g->ny,
                "row strip [%d,%d) outside grid of %d rows", g->row_begin,
                g->row_begin + g->row_count, g->ny);
    const bool full = g->row_begin == 0 && g->row_count == g->ny;
    LDC_REQUIRE(full || g->batch == 1, "row strips cannot be combined with batch > 1");
    LDC_REQUIRE(!need_full_grid || full, "this entry point needs the full grid (no row strip)");
    LDC_REQUIRE(g->dx > 0.0 && g->dy > 0.0, "dx, dy must be positive");
    LDC_REQUIRE(g->dt_dev != nullptr || g->dt > 0.0, "dt must be positive");
    return LDC_OK;
}

int sm_count()
{
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;  // B200
    }
    return cached;
}

}  // namespace ldc

extern "C" const char *ldc_last_error(void) { return ldc::g_error; }

extern "C" int ldc_version(void) { return 100; }

extern "C" int ldc_device_info(int32_t *sm, int32_t *cc_major, int32_t *cc_minor, int64_t *l2_bytes)
{
    int dev = 0;
    LDC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    LDC_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm) *sm = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = prop.l2CacheSize;
    return LDC_OK;
}
