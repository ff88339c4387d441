// dist.cu -- NCCL plumbing for the row-partitioned eigensolve (BASELINE.json config 5): one process per GPU,
// the Lanczos vector is exchanged by ncclAllGather over NVLink and the recurrence scalars by ncclAllReduce.
// libnccl.so.2 is dlopen'ed on first use (the copy torch already loaded when the host process is Python), so the
// single-GPU library has no link-time NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "bh_internal.h"

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

const char* load_nccl()
{
    if (g_nccl.handle) return nullptr;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return "libnccl.so.2 not found";
#define SYM(field, name)                                              \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
    if (!g_nccl.field) return "missing NCCL symbol " name;
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return nullptr;
}
}  // namespace

#define BH_NCCL(ctx, expr)                                                                           \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != ncclSuccess)                                                                       \
            return bh_fail((ctx), BH_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

extern "C" int bh_dist_unique_id(void* id128)
{
    if (!id128) return BH_ERR_ARG;
    if (const char* e = load_nccl()) return bh_fail(nullptr, BH_ERR_CUDA, e);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return bh_fail(nullptr, BH_ERR_CUDA, "ncclGetUniqueId failed");
    std::memcpy(id128, &id, sizeof(id));
    return BH_OK;
}

extern "C" int bh_dist_init(bh_ctx* ctx, int world, int rank, const void* id128)
{
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return bh_fail(ctx, BH_ERR_ARG, "bh_dist_init: bad argument");
    if (const char* e = load_nccl()) return bh_fail(ctx, BH_ERR_CUDA, e);
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    BH_NCCL(ctx, g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->world = world;
    ctx->rank = rank;
    return BH_OK;
}

extern "C" int bh_dist_finalize(bh_ctx* ctx)
{
    if (!ctx) return BH_ERR_ARG;
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
        ctx->nccl_comm = nullptr;
    }
    ctx->world = 1;
    ctx->rank = 0;
    return BH_OK;
}

// The collectives below are issued only by a context that is row-partitioned (bh_setup_partitioned): a context that called
// bh_dist_init and then a plain bh_setup solves its own grid points and must not meet its peers in a collective.
int bh_dist_allreduce_sum(bh_ctx* ctx, double* buf, int64_t count)
{
    if (ctx->world < 2 || !ctx->partitioned) return BH_OK;
    BH_NCCL(ctx, g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    return BH_OK;
}

int bh_dist_allreduce_max(bh_ctx* ctx, double* buf, int64_t count)
{
    if (ctx->world < 2 || !ctx->partitioned) return BH_OK;
    BH_NCCL(ctx, g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclMax, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    return BH_OK;
}

int bh_dist_allgather(bh_ctx* ctx, const double* send, double* recv, int64_t count_per_rank)
{
    if (ctx->world < 2 || !ctx->partitioned) {
        BH_CUDA(ctx, cudaMemcpyAsync(recv, send, sizeof(double) * count_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return BH_OK;
    }
    BH_NCCL(ctx, g_nccl.AllGather(send, recv, (size_t)count_per_rank, ncclDouble, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    return BH_OK;
}
