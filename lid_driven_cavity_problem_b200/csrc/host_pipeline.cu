// ldc_residual_host: the residual plug point with HOST buffers in and out -- what the reference's
// callers hold (PETSc's callback and scipy's fsolve pass numpy arrays,
// petsc_solver_wrapper.py:42-48, scipy_solver_wrapper.py:35; the native residual
// c++/residual_function.cpp:10-43 takes and returns host arrays).
//
// A 1024^2 evaluation moves 42 MB in and 25 MB out for a 14 us kernel, so the call is a copy
// pipeline, not a kernel launch.  The grid is cut into row blocks; for block b
//   host threads  : pageable X / U_old / V_old rows -> pinned staging          (memcpy, parallel)
//   stream `in`   : staging -> HBM                                             (H2D DMA)
//   stream `run`  : residual kernel on the block's rows, once block b+1's X rows (its upper halo)
//                   have arrived                                               (ldc_residual, strip form)
//   stream `out`  : the block's F rows -> pinned staging                       (D2H DMA, full duplex
//                                                                               with the H2D of later blocks)
//   host threads  : staging -> the caller's F                                  (memcpy, parallel; skipped
//                                                                               when F is page-locked already)
// so the PCIe link is busy in both directions and the host copies hide behind the DMAs.
// Staging buffers, device buffers, streams, events and the copy threads are created once per
// process and reused (grow-only); calls are serialised by a mutex.
#include <emmintrin.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "ldc_internal.cuh"

namespace ldc {
namespace {

// ---- a small persistent pool for the host-side copies (torchrun pins OMP_NUM_THREADS to 1 for
// its workers, so the library brings its own threads)
class CopyPool {
public:
    explicit CopyPool(int n) : n_(n), stop_(false), epoch_(0), pending_(0)
    {
        for (int t = 0; t < n_ - 1; ++t) workers_.emplace_back([this, t] { loop(t + 1); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    // dst[0..bytes) = src[0..bytes), split over the pool; returns when done
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (bytes < (1u << 20) || n_ == 1) {
            stream_copy(static_cast<char *>(dst), static_cast<const char *>(src), bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = static_cast<char *>(dst);
            src_ = static_cast<const char *>(src);
            bytes_ = bytes;
            pending_ = n_ - 1;
            ++epoch_;
        }
        cv_.notify_all();
        slice(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
    int threads() const { return n_; }

private:
    void slice(int t)
    {
        const size_t per = ((bytes_ + n_ - 1) / n_ + 4095) & ~size_t(4095);
        const size_t a = per * t;
        if (a >= bytes_) return;
        const size_t len = (a + per <= bytes_) ? per : bytes_ - a;
        stream_copy(dst_ + a, src_ + a, len);
    }
    // memcpy with non-temporal stores: the destination is either a pinned staging buffer the DMA
    // engine reads next or the caller's fresh result array -- neither is read by this core again,
    // so write-allocating it into the cache only doubles the memory traffic of the copy
    static void stream_copy(char *dst, const char *src, size_t len)
    {
        if (len < (64u << 10) || (reinterpret_cast<uintptr_t>(dst) & 15) != 0) {
            memcpy(dst, src, len);
            return;
        }
        const size_t body = len & ~size_t(63);
        for (size_t k = 0; k < body; k += 64) {
            const __m128i a0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + k));
            const __m128i a1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + k + 16));
            const __m128i a2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + k + 32));
            const __m128i a3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + k + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + k), a0);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + k + 16), a1);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + k + 32), a2);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + k + 48), a3);
        }
        _mm_sfence();
        if (body < len) memcpy(dst + body, src + body, len - body);
    }
    void loop(int t)
    {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            slice(t);
            std::lock_guard<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    const int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    bool stop_;
    unsigned long epoch_;
    int pending_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0;
};

constexpr int kMaxBlocks = 8;

struct HostPipeline {
    std::mutex lock;
    CopyPool *pool = nullptr;
    int device = -1;
    size_t cap_doubles = 0;   // capacity in cells*3 doubles of the X / F buffers
    double *pin_in = nullptr, *pin_out = nullptr;      // pinned: [X | U_old | V_old], [F]
    double *dX = nullptr, *dU = nullptr, *dV = nullptr, *dF = nullptr;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kMaxBlocks] = {}, ev_run[kMaxBlocks] = {}, ev_out[kMaxBlocks] = {};

    int ensure(size_t n3)   // n3 = 3*nx*ny doubles
    {
        int dev = 0;
        LDC_CUDA(cudaGetDevice(&dev));
        if (!pool) {
            int nt = (int)std::thread::hardware_concurrency();
            if (const char *lw = getenv("LOCAL_WORLD_SIZE")) {
                const int ranks = atoi(lw);
                if (ranks > 1) nt /= ranks;
            }
            if (const char *e = getenv("LDC_HOST_COPY_THREADS")) nt = atoi(e);
            if (nt > 8) nt = 8;   // measured on the B200 box (16 cores): 4 threads 1.8 ms, 8 threads 1.43 ms, 12-16 no better
            if (nt < 1) nt = 1;
            pool = new CopyPool(nt);
        }
        if (dev != device) {   // first use, or the caller switched devices: start over on this one
            release();
            device = dev;
            LDC_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
            LDC_CUDA(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
            LDC_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
            for (int b = 0; b < kMaxBlocks; ++b) {
                LDC_CUDA(cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming));
                LDC_CUDA(cudaEventCreateWithFlags(&ev_run[b], cudaEventDisableTiming));
                LDC_CUDA(cudaEventCreateWithFlags(&ev_out[b], cudaEventDisableTiming));
            }
        }
        if (n3 > cap_doubles) {
            free_buffers();
            const size_t in_doubles = n3 + 2 * (n3 / 3) + 64;   // X + U_old + V_old (each <= nx*ny)
            LDC_CUDA(cudaHostAlloc(&pin_in, in_doubles * sizeof(double), cudaHostAllocDefault));
            LDC_CUDA(cudaHostAlloc(&pin_out, n3 * sizeof(double), cudaHostAllocDefault));
            LDC_CUDA(cudaMalloc(&dX, n3 * sizeof(double)));
            LDC_CUDA(cudaMalloc(&dF, n3 * sizeof(double)));
            LDC_CUDA(cudaMalloc(&dU, (n3 / 3 + 16) * sizeof(double)));
            LDC_CUDA(cudaMalloc(&dV, (n3 / 3 + 16) * sizeof(double)));
            cap_doubles = n3;
        }
        return LDC_OK;
    }
    void free_buffers()
    {
        if (pin_in) cudaFreeHost(pin_in);
        if (pin_out) cudaFreeHost(pin_out);
        if (dX) cudaFree(dX);
        if (dF) cudaFree(dF);
        if (dU) cudaFree(dU);
        if (dV) cudaFree(dV);
        pin_in = pin_out = dX = dF = dU = dV = nullptr;
        cap_doubles = 0;
    }
    void release()
    {
        free_buffers();
        for (int b = 0; b < kMaxBlocks; ++b) {
            if (ev_in[b]) cudaEventDestroy(ev_in[b]);
            if (ev_run[b]) cudaEventDestroy(ev_run[b]);
            if (ev_out[b]) cudaEventDestroy(ev_out[b]);
            ev_in[b] = ev_run[b] = ev_out[b] = nullptr;
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_out) cudaStreamDestroy(s_out);
        s_in = s_run = s_out = nullptr;
    }
};

HostPipeline g_pipe;

}  // namespace
}  // namespace ldc

using namespace ldc;

extern "C" int ldc_residual_host(const ldc_grid *g, const double *X_host, const double *U_old_host,
                                 const double *V_old_host, double *F_host, int32_t variant)
{
    if (int rc = validate_grid(g, true)) return rc;
    LDC_REQUIRE(g->batch == 1, "ldc_residual_host: one cavity per call");
    LDC_REQUIRE(g->dt_dev == nullptr && g->bc_dev == nullptr, "ldc_residual_host: dt / bc are host scalars");
    LDC_REQUIRE(X_host && U_old_host && V_old_host && F_host, "ldc_residual_host: null host pointer");
    LDC_REQUIRE(X_host != F_host, "ldc_residual_host: F must not alias X");
    const long nx = g->nx, ny = g->ny;
    const size_t n3 = (size_t)3 * nx * ny;
    // A result buffer that is already page-locked (the Python forwarder hands out numpy arrays backed
    // by torch's pinned-memory cache) receives the D2H copies directly: no staging, no copy-out.
    bool f_pinned = false;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, F_host) == cudaSuccess)
            f_pinned = attr.type == cudaMemoryTypeHost;
        else
            cudaGetLastError();   // pageable memory is not an error here
    }
    HostPipeline &p = g_pipe;
    std::lock_guard<std::mutex> guard(p.lock);
    // LDC_HOST_PROFILE=1: host-clock marks of the pipeline phases on stderr (us since call entry)
    static const bool profile = getenv("LDC_HOST_PROFILE") != nullptr;
    const auto t_entry = std::chrono::steady_clock::now();
    auto mark = [&](const char *what, int b) {
        if (profile)
            fprintf(stderr, "[ldc_residual_host] %8.1f us  %s %d\n",
                    std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_entry).count(), what, b);
    };
    if (int rc = p.ensure(n3)) return rc;
    mark("buffers ready", 0);

    // row blocks: enough of them to overlap the two DMA directions, each at least ~2 MB of X
    int nb = (int)(n3 * sizeof(double) / (6u << 20));
    if (nb > kMaxBlocks) nb = kMaxBlocks;
    if (nb > ny / 2) nb = (int)(ny / 2);
    if (nb < 1) nb = 1;
    // block boundaries on even rows: keeps every block's U_old / V_old / X pointers 16-byte aligned
    // for even nx, which the default kernel (bulk-copy rings, 128-bit accesses) requires
    auto row_of = [&](int b) { return b >= nb ? ny : (long)(((ny * (long)b) / nb) & ~1L); };
    double *pinX = p.pin_in, *pinU = p.pin_in + n3, *pinV = pinU + (size_t)(nx - 1) * ny + 8;
    pinV = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(pinV) + 63) & ~uintptr_t(63));

    auto launch_block = [&](int b) -> int {   // kernel + D2H of block b (its inputs are resident)
        const long r0 = row_of(b), r1 = row_of(b + 1);
        ldc_grid sub = *g;
        sub.row_begin = (int32_t)r0;
        sub.row_count = (int32_t)(r1 - r0);
        LDC_CUDA(cudaStreamWaitEvent(p.s_run, p.ev_in[b + 1 < nb ? b + 1 : b], 0));
        if (int rc = ldc_residual(&sub, p.dX + 3 * nx * r0, p.dU + (nx - 1) * r0, p.dV + nx * r0, p.dF + 3 * nx * r0,
                                  nullptr, nullptr, variant, p.s_run))
            return rc;
        LDC_CUDA(cudaEventRecord(p.ev_run[b], p.s_run));
        LDC_CUDA(cudaStreamWaitEvent(p.s_out, p.ev_run[b], 0));
        LDC_CUDA(cudaMemcpyAsync((f_pinned ? F_host : p.pin_out) + 3 * nx * r0, p.dF + 3 * nx * r0,
                                 sizeof(double) * 3 * nx * (r1 - r0), cudaMemcpyDeviceToHost, p.s_out));
        LDC_CUDA(cudaEventRecord(p.ev_out[b], p.s_out));
        return LDC_OK;
    };

    for (int b = 0; b < nb; ++b) {
        const long r0 = row_of(b), r1 = row_of(b + 1);
        const long v1 = r1 < ny - 1 ? r1 : ny - 1;   // V_old has ny-1 rows
        // stage and send the block's rows: X first (the previous block's kernel waits for them)
        p.pool->copy(pinX + 3 * nx * r0, X_host + 3 * nx * r0, sizeof(double) * 3 * nx * (r1 - r0));
        LDC_CUDA(cudaMemcpyAsync(p.dX + 3 * nx * r0, pinX + 3 * nx * r0, sizeof(double) * 3 * nx * (r1 - r0),
                                 cudaMemcpyHostToDevice, p.s_in));
        p.pool->copy(pinU + (nx - 1) * r0, U_old_host + (nx - 1) * r0, sizeof(double) * (nx - 1) * (r1 - r0));
        LDC_CUDA(cudaMemcpyAsync(p.dU + (nx - 1) * r0, pinU + (nx - 1) * r0, sizeof(double) * (nx - 1) * (r1 - r0),
                                 cudaMemcpyHostToDevice, p.s_in));
        if (v1 > r0) {
            p.pool->copy(pinV + nx * r0, V_old_host + nx * r0, sizeof(double) * nx * (v1 - r0));
            LDC_CUDA(cudaMemcpyAsync(p.dV + nx * r0, pinV + nx * r0, sizeof(double) * nx * (v1 - r0),
                                     cudaMemcpyHostToDevice, p.s_in));
        }
        LDC_CUDA(cudaEventRecord(p.ev_in[b], p.s_in));
        mark("staged + H2D enqueued, block", b);
        // block b-1 has everything now: its own rows (ev_in[b-1], earlier on the same stream) and
        // its upper halo row, the first row of block b
        if (b >= 1)
            if (int rc = launch_block(b - 1)) return rc;
    }
    if (int rc = launch_block(nb - 1)) return rc;

    // drain: copy each block out of the staging buffer as soon as its DMA has landed
    for (int b = 0; b < nb; ++b) {
        const long r0 = row_of(b), r1 = row_of(b + 1);
        LDC_CUDA(cudaEventSynchronize(p.ev_out[b]));
        mark("D2H landed, block", b);
        if (!f_pinned)
            p.pool->copy(F_host + 3 * nx * r0, p.pin_out + 3 * nx * r0, sizeof(double) * 3 * nx * (r1 - r0));
    }
    mark("copied out", nb);
    return LDC_OK;
}
