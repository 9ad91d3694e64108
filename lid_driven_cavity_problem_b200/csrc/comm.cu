// NCCL transport of the row-strip decomposition (libldc_b200_comm.so, see
// include/ldc_b200_comm.h).  Halo rows are latency-bound messages (3*nx doubles = 96 KiB at
// nx = 4096): both directions go into ONE NCCL group so they share a launch, and the calls are
// made from C (a few microseconds of host time) rather than through a Python collective layer,
// which matters because the kernels they feed run for only tens of microseconds.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/ldc_b200_comm.h"

struct PeerRegion {
    void *local, *prev, *next;
};

struct ldc_comm {
    ncclComm_t comm;
    int rank, nranks;
    ldc_comm_ops ops;
    std::vector<PeerRegion> regions;
};

static thread_local char g_comm_error[512] = "";

static int comm_fail(const char *what, const char *detail)
{
    snprintf(g_comm_error, sizeof(g_comm_error), "%s: %s", what, detail);
    return LDC_ERR_CUDA;
}

#define LDC_RT(call)                                                         \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) {                                            \
            (void)cudaGetLastError(); /* do not leave it for the next launch check */ \
            return comm_fail(#call, cudaGetErrorString(e__));                \
        }                                                                    \
    } while (0)

#define LDC_NCCL(call)                                                       \
    do {                                                                     \
        ncclResult_t r__ = (call);                                           \
        if (r__ != ncclSuccess) return comm_fail(#call, ncclGetErrorString(r__)); \
    } while (0)

static int halo_cb(void *ctx, double *first, int64_t row_doubles, int64_t row_count, void *stream)
{
    return ldc_comm_halo_exchange((ldc_comm *)ctx, first, row_doubles, row_count, stream);
}

static int allreduce_cb(void *ctx, double *buf, int64_t count, int32_t op, void *stream)
{
    return ldc_comm_allreduce((ldc_comm *)ctx, buf, count, op, stream);
}

static int allgather_cb(void *ctx, double *buf, int64_t count, void *stream)
{
    return ldc_comm_allgather((ldc_comm *)ctx, buf, count, stream);
}

extern "C" const char *ldc_comm_last_error(void) { return g_comm_error; }

extern "C" int ldc_comm_unique_id(void *id_out)
{
    static_assert(sizeof(ncclUniqueId) <= LDC_COMM_ID_BYTES, "ncclUniqueId larger than expected");
    if (!id_out) return comm_fail("ldc_comm_unique_id", "null argument");
    ncclUniqueId id;
    LDC_NCCL(ncclGetUniqueId(&id));
    memset(id_out, 0, LDC_COMM_ID_BYTES);
    memcpy(id_out, &id, sizeof(id));
    return LDC_OK;
}

extern "C" int ldc_comm_create(const void *id_bytes, int32_t nranks, int32_t rank, ldc_comm **out)
{
    if (!id_bytes || !out || nranks < 1 || rank < 0 || rank >= nranks)
        return comm_fail("ldc_comm_create", "bad arguments");
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ldc_comm *c = new ldc_comm();
    c->rank = rank;
    c->nranks = nranks;
    ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return comm_fail("ncclCommInitRank", ncclGetErrorString(r));
    }
    c->ops.ctx = c;
    c->ops.rank = rank;
    c->ops.nranks = nranks;
    c->ops.halo_exchange = halo_cb;
    c->ops.allreduce = allreduce_cb;
    c->ops.allgather = allgather_cb;
    *out = c;
    return LDC_OK;
}

extern "C" void ldc_comm_destroy(ldc_comm *c)
{
    if (!c) return;
    // regions the caller did not hand back: unmap the neighbours', free ours (no collective here)
    for (const PeerRegion &r : c->regions) {
        if (r.prev) cudaIpcCloseMemHandle(r.prev);
        if (r.next) cudaIpcCloseMemHandle(r.next);
        cudaFree(r.local);
    }
    c->regions.clear();
    ncclCommDestroy(c->comm);
    delete c;
}

extern "C" const ldc_comm_ops *ldc_comm_get_ops(ldc_comm *c) { return c ? &c->ops : nullptr; }

extern "C" int ldc_comm_halo_exchange(ldc_comm *c, double *first, int64_t row_doubles,
                                      int64_t row_count, void *stream)
{
    if (!c || !first || row_doubles <= 0 || row_count <= 0)
        return comm_fail("ldc_comm_halo_exchange", "bad arguments");
    if (c->nranks == 1) return LDC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double *ghost_below = first - row_doubles;
    double *last = first + (row_count - 1) * row_doubles;
    double *ghost_above = first + row_count * row_doubles;
    LDC_NCCL(ncclGroupStart());
    if (c->rank > 0) {
        LDC_NCCL(ncclSend(first, (size_t)row_doubles, ncclDouble, c->rank - 1, c->comm, st));
        LDC_NCCL(ncclRecv(ghost_below, (size_t)row_doubles, ncclDouble, c->rank - 1, c->comm, st));
    }
    if (c->rank < c->nranks - 1) {
        LDC_NCCL(ncclSend(last, (size_t)row_doubles, ncclDouble, c->rank + 1, c->comm, st));
        LDC_NCCL(ncclRecv(ghost_above, (size_t)row_doubles, ncclDouble, c->rank + 1, c->comm, st));
    }
    LDC_NCCL(ncclGroupEnd());
    return LDC_OK;
}

extern "C" int ldc_comm_allreduce(ldc_comm *c, double *buf, int64_t count, int32_t op, void *stream)
{
    if (!c || !buf || count <= 0) return comm_fail("ldc_comm_allreduce", "bad arguments");
    if (c->nranks == 1) return LDC_OK;
    LDC_NCCL(ncclAllReduce(buf, buf, (size_t)count, ncclDouble, op == 1 ? ncclMax : ncclSum, c->comm,
                           (cudaStream_t)stream));
    return LDC_OK;
}

extern "C" int ldc_comm_allgather(ldc_comm *c, double *buf, int64_t count, void *stream)
{
    if (!c || !buf || count <= 0) return comm_fail("ldc_comm_allgather", "bad arguments");
    if (c->nranks == 1) return LDC_OK;
    LDC_NCCL(ncclAllGather(buf + (size_t)c->rank * count, buf, (size_t)count, ncclDouble, c->comm,
                           (cudaStream_t)stream));
    return LDC_OK;
}

// ---- peer-mapped regions: the memory behind ldc_residual_peer (include/ldc_b200.h).  Every rank
// allocates the same number of bytes; CUDA IPC handles travel through one NCCL all-gather and the
// two neighbours' regions are mapped into this process, so kernels can load the neighbour's
// boundary rows and store its flag words directly over NVLink.
extern "C" int ldc_comm_peer_alloc(ldc_comm *c, int64_t bytes, void **local_out, void **prev_out,
                                   void **next_out)
{
    if (!c || bytes <= 0 || !local_out || !prev_out || !next_out)
        return comm_fail("ldc_comm_peer_alloc", "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PeerRegion r{nullptr, nullptr, nullptr};
    LDC_RT(cudaMalloc(&r.local, (size_t)bytes));
    LDC_RT(cudaMemset(r.local, 0, (size_t)bytes));
    if (c->nranks > 1) {
        cudaIpcMemHandle_t mine;
        LDC_RT(cudaIpcGetMemHandle(&mine, r.local));
        std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
        void *hbuf = nullptr;
        LDC_RT(cudaMalloc(&hbuf, 64 * (size_t)c->nranks));
        LDC_RT(cudaMemcpy((char *)hbuf + 64 * (size_t)c->rank, &mine, 64, cudaMemcpyHostToDevice));
        LDC_NCCL(ncclAllGather((char *)hbuf + 64 * (size_t)c->rank, hbuf, 64, ncclChar, c->comm, 0));
        LDC_RT(cudaStreamSynchronize(0));
        LDC_RT(cudaMemcpy(all.data(), hbuf, 64 * (size_t)c->nranks, cudaMemcpyDeviceToHost));
        LDC_RT(cudaFree(hbuf));
        if (c->rank > 0)
            LDC_RT(cudaIpcOpenMemHandle(&r.prev, all[c->rank - 1], cudaIpcMemLazyEnablePeerAccess));
        if (c->rank < c->nranks - 1)
            LDC_RT(cudaIpcOpenMemHandle(&r.next, all[c->rank + 1], cudaIpcMemLazyEnablePeerAccess));
    }
    c->regions.push_back(r);
    *local_out = r.local;
    *prev_out = r.prev;
    *next_out = r.next;
    return LDC_OK;
}

extern "C" int ldc_comm_peer_free(ldc_comm *c, void *local)
{
    if (!c || !local) return comm_fail("ldc_comm_peer_free", "bad arguments");
    for (size_t k = 0; k < c->regions.size(); ++k) {
        if (c->regions[k].local != local) continue;
        PeerRegion r = c->regions[k];
        c->regions.erase(c->regions.begin() + (long)k);
        LDC_RT(cudaDeviceSynchronize());
        if (r.prev) LDC_RT(cudaIpcCloseMemHandle(r.prev));
        if (r.next) LDC_RT(cudaIpcCloseMemHandle(r.next));
        if (c->nranks > 1) {
            // nobody frees before every neighbour has dropped its mapping
            double *tok = nullptr;
            LDC_RT(cudaMalloc((void **)&tok, sizeof(double)));
            LDC_RT(cudaMemset(tok, 0, sizeof(double)));
            LDC_NCCL(ncclAllReduce(tok, tok, 1, ncclDouble, ncclSum, c->comm, 0));
            LDC_RT(cudaStreamSynchronize(0));
            LDC_RT(cudaFree(tok));
        }
        LDC_RT(cudaFree(r.local));
        return LDC_OK;
    }
    return comm_fail("ldc_comm_peer_free", "unknown region");
}
