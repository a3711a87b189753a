// Multi-GPU plumbing: one process per GPU, NCCL communicator owned by the context.
// NCCL is resolved with dlopen at first use (the copy already loaded by the host process, e.g.
// torch's, wins), so libvrec.so has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "vrec_internal.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.handle) return VREC_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        vrec_set_error("cannot load libnccl.so.2: %s", dlerror());
        return VREC_ENCCL;
    }
#define VREC_SYM(field, name)                                         \
    *(void **)(&g_nccl.field) = dlsym(h, name);                       \
    if (!g_nccl.field) {                                              \
        vrec_set_error("libnccl lacks %s", name);                     \
        return VREC_ENCCL;                                            \
    }
    VREC_SYM(GetUniqueId, "ncclGetUniqueId")
    VREC_SYM(CommInitRank, "ncclCommInitRank")
    VREC_SYM(CommDestroy, "ncclCommDestroy")
    VREC_SYM(AllGather, "ncclAllGather")
    VREC_SYM(AllReduce, "ncclAllReduce")
    VREC_SYM(GetErrorString, "ncclGetErrorString")
#undef VREC_SYM
    g_nccl.handle = h;
    return VREC_OK;
}

int nccl_check(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return VREC_OK;
    vrec_set_error("%s -> %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
    return VREC_ENCCL;
}

}  // namespace

extern "C" int vrec_comm_unique_id(void *out128) {
    if (!out128) return VREC_EINVAL;
    VREC_TRY(nccl_load());
    ncclUniqueId id;
    VREC_TRY(nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return VREC_OK;
}

extern "C" int vrec_comm_init(vrec_ctx *ctx, int rank, int world, const void *unique_id128) {
    if (!ctx || !unique_id128 || world < 1 || rank < 0 || rank >= world) {
        vrec_set_error("vrec_comm_init: bad argument");
        return VREC_EINVAL;
    }
    VREC_TRY(nccl_load());
    VREC_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id128, 128);
    ncclComm_t comm = nullptr;
    VREC_TRY(nccl_check(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank"));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return VREC_OK;
}

extern "C" int vrec_comm_rank(vrec_ctx *ctx) { return ctx ? ctx->rank : 0; }
extern "C" int vrec_comm_world(vrec_ctx *ctx) { return ctx ? ctx->world : 1; }

void vrec_comm_destroy(vrec_ctx *ctx) {
    if (ctx && ctx->comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
    }
}

// all-gather of `bytes` bytes per rank (device buffers); used once per graph to exchange CUDA IPC handles --
// the iteration itself exchanges x' with peer stores from the sweep kernel, not with a collective
int vrec_comm_allgather_bytes(vrec_ctx *ctx, const void *send, void *recv, size_t bytes) {
    if (ctx->world <= 1) {
        cudaError_t e = cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) {
            vrec_set_error("cudaMemcpyAsync -> %s", cudaGetErrorString(e));
            return VREC_ECUDA;
        }
        return VREC_OK;
    }
    if (!ctx->comm) {
        vrec_set_error("vrec_comm_init has not been called on this context");
        return VREC_ENCCL;
    }
    return nccl_check(g_nccl.AllGather(send, recv, bytes, ncclChar, (ncclComm_t)ctx->comm, ctx->stream), "ncclAllGather");
}
