// nccl_dl.cpp — NCCL loaded at run time (dlopen) for the data-parallel mode.
//
// One process per GPU; the host distributes the 128-byte unique id (torch.distributed
// broadcast in the Python layer, any transport in Julia).  The only collectives on the path are
// the per-minibatch gradient sum (P floats) and the per-epoch loss-history sum (doubles).
// dlopen instead of link-time dependency: inside a process that already loaded torch's bundled
// libnccl.so.2 the same soname resolves to that copy (no second NCCL in the address space);
// stand-alone (Julia) it resolves to the system library.  PPO_B200_NCCL_LIB overrides the path.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ppo {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_api;

int load_api() {
    if (g_api.handle) return PPO_OK;
    const char* override_path = getenv("PPO_B200_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        if (!nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror());
        return PPO_ERR_NCCL;
    }
#define PPO_SYM(field, name)                                              \
    *(void**)(&g_api.field) = dlsym(h, name);                             \
    if (!g_api.field) { set_error("NCCL: missing symbol %s", name); return PPO_ERR_NCCL; }
    PPO_SYM(GetUniqueId, "ncclGetUniqueId");
    PPO_SYM(CommInitRank, "ncclCommInitRank");
    PPO_SYM(CommDestroy, "ncclCommDestroy");
    PPO_SYM(AllReduce, "ncclAllReduce");
    PPO_SYM(GetErrorString, "ncclGetErrorString");
#undef PPO_SYM
    g_api.handle = h;
    return PPO_OK;
}

#define PPO_NCCL(expr)                                                                         \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                              \
        if (_r != ncclSuccess) {                                                               \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_api.GetErrorString(_r)); \
            return PPO_ERR_NCCL;                                                               \
        }                                                                                      \
    } while (0)

}  // namespace

int nccl_unique_id(void* id128) {
    PPO_TRY(load_api());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    PPO_NCCL(g_api.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return PPO_OK;
}

int nccl_init(ppo_ctx* ctx, int nranks, int rank, const void* id128) {
    PPO_TRY(load_api());
    PPO_REQUIRE(ctx->nccl_comm == nullptr, "communicator already initialised");
    PPO_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad rank %d of %d", rank, nranks);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    PPO_NCCL(g_api.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = (void*)comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return PPO_OK;
}

int nccl_destroy(ppo_ctx* ctx) {
    if (ctx->nccl_comm) {
        PPO_NCCL(g_api.CommDestroy((ncclComm_t)ctx->nccl_comm));
        ctx->nccl_comm = nullptr;
        ctx->nranks = 1;
        ctx->rank = 0;
    }
    return PPO_OK;
}

int nccl_allreduce_f32(ppo_ctx* ctx, float* d_buf, int64_t n) {
    PPO_REQUIRE(ctx->nccl_comm != nullptr, "no communicator");
    PPO_NCCL(g_api.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return PPO_OK;
}

int nccl_allreduce_f32_on(ppo_ctx* ctx, float* d_buf, int64_t n, cudaStream_t stream) {
    PPO_REQUIRE(ctx->nccl_comm != nullptr, "no communicator");
    PPO_NCCL(g_api.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, stream));
    return PPO_OK;
}

bool dp_overlap(ppo_ctx* ctx) {
    static const int off = getenv("PPO_B200_NO_DP_OVERLAP") ? atoi(getenv("PPO_B200_NO_DP_OVERLAP")) : 0;
    return ctx->nccl_comm != nullptr && ctx->nranks > 1 && !off && !ctx->p2p_grads;
}

int grads_ready(ppo_ctx* ctx, float* d_slice, int64_t n) {
    if (!dp_overlap(ctx) || n <= 0) return PPO_OK;
    if (ctx->comm_stream == nullptr) {
        PPO_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
        PPO_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        PPO_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    PPO_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    PPO_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
    PPO_TRY(nccl_allreduce_f32_on(ctx, d_slice, n, ctx->comm_stream));
    ctx->comm_pending = true;
    return PPO_OK;
}

int grads_join(ppo_ctx* ctx) {
    if (!ctx->comm_pending) return PPO_OK;
    PPO_CUDA(cudaEventRecord(ctx->ev_join, ctx->comm_stream));
    PPO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    ctx->comm_pending = false;
    return PPO_OK;
}

int nccl_allreduce_f64(ppo_ctx* ctx, double* d_buf, int64_t n) {
    PPO_REQUIRE(ctx->nccl_comm != nullptr, "no communicator");
    PPO_NCCL(g_api.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return PPO_OK;
}

}  // namespace ppo
