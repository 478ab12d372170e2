// Data-parallel gradient exchange.  The reference is single-process; this is the one
// exchange step the sharded path needs: a sum-allreduce of the flat gradient bucket after backward.
// NCCL is bound at run time (dlopen) so the library loads on boxes without it and shares the copy
// a launcher such as torch.distributed has already mapped.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    if (a.handle || a.ok) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) return a;
#define TP_SYM(field, sym) a.field = (decltype(a.field))dlsym(a.handle, sym)
    TP_SYM(GetUniqueId, "ncclGetUniqueId");
    TP_SYM(CommInitRank, "ncclCommInitRank");
    TP_SYM(CommDestroy, "ncclCommDestroy");
    TP_SYM(AllReduce, "ncclAllReduce");
    TP_SYM(Broadcast, "ncclBroadcast");
    TP_SYM(GetErrorString, "ncclGetErrorString");
#undef TP_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Broadcast && a.GetErrorString;
    return a;
}

#define TP_NCCL(expr)                                                                         \
    do {                                                                                      \
        ncclResult_t _r = (expr);                                                             \
        if (_r != ncclSuccess) {                                                              \
            tp::set_error("%s failed: %s", #expr, api().GetErrorString(_r));                  \
            return TP_ERR_COMM;                                                               \
        }                                                                                     \
    } while (0)

int need_api() {
    if (!api().ok) {
        tp::set_error("NCCL is not available: dlopen(libnccl.so.2) failed or symbols missing (%s)", dlerror());
        return TP_ERR_COMM;
    }
    return TP_OK;
}

}  // namespace

extern "C" {

int tp_comm_unique_id(void* out128) {
    TP_CHECK_ARG(out128, "tp_comm_unique_id: NULL out pointer");
    int rc = need_api();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    TP_NCCL(api().GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return TP_OK;
}

int tp_comm_init(tp_ctx* ctx, int rank, int world, const void* unique_id128) {
    TP_CHECK_ARG(ctx && unique_id128 && world >= 1 && rank >= 0 && rank < world, "tp_comm_init: bad arguments");
    if (ctx->nccl_comm) {
        // one communicator per context: a second trainer on the same thread joins the one that exists
        TP_CHECK_ARG(ctx->rank == rank && ctx->world == world, "tp_comm_init: communicator already initialised as rank %d of %d", ctx->rank, ctx->world);
        return TP_OK;
    }
    int rc = need_api();
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof id);
    ncclComm_t comm;
    TP_NCCL(api().CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return TP_OK;
}

int tp_comm_destroy(tp_ctx* ctx) {
    if (ctx && ctx->nccl_comm && api().ok) {
        api().CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return TP_OK;
}

int tp_allreduce_sum(tp_ctx* ctx, tp_buf* buf, size_t n) {
    TP_CHECK_ARG(ctx, "tp_allreduce_sum: NULL ctx");
    TP_NEED(buf, n, "buf");
    if (ctx->world == 1 && !ctx->nccl_comm) return TP_OK;       // single replica: the sum is the buffer itself
    TP_CHECK_ARG(ctx->nccl_comm, "tp_allreduce_sum: tp_comm_init has not been called");
    TP_NCCL(api().AllReduce(buf->ptr, buf->ptr, n, ncclFloat, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ctx->launches++;
    return TP_OK;
}

int tp_broadcast(tp_ctx* ctx, tp_buf* buf, size_t n, int root) {
    TP_CHECK_ARG(ctx, "tp_broadcast: NULL ctx");
    TP_NEED(buf, n, "buf");
    if (ctx->world == 1 && !ctx->nccl_comm) return TP_OK;
    TP_CHECK_ARG(ctx->nccl_comm, "tp_broadcast: tp_comm_init has not been called");
    TP_NCCL(api().Broadcast(buf->ptr, buf->ptr, n, ncclFloat, root, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ctx->launches++;
    return TP_OK;
}

}  // extern "C"
