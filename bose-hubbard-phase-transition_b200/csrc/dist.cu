// dist.cu -- NCCL plumbing for the row-partitioned eigensolve (BASELINE.json config 5): one process per GPU,
// the Lanczos vector is exchanged by ncclAllGather over NVLink and the recurrence scalars by ncclAllReduce.
// libnccl.so.2 is dlopen'ed on first use (the copy torch already loaded when the host process is Python), so the
// single-GPU library has no link-time NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bh_internal.h"
#include "halo_plan.h"

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
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
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    g_nccl.CommSplit = reinterpret_cast<decltype(g_nccl.CommSplit)>(dlsym(h, "ncclCommSplit"));  // NCCL >= 2.18; optional
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

void bh_dist_release_halo(bh_ctx* ctx)
{
    if (ctx->d_rem_ptr) cudaFree(ctx->d_rem_ptr);
    if (ctx->d_rem_col) cudaFree(ctx->d_rem_col);
    if (ctx->d_rem_amp) cudaFree(ctx->d_rem_amp);
    ctx->d_rem_ptr = ctx->d_rem_col = nullptr;
    ctx->d_rem_amp = nullptr;
    ctx->rem_nnz = 0;
    ctx->halo_ready = false;
    ctx->halo_send.clear();
    ctx->halo_recv.clear();
    ctx->halo_recv_elems = 0;
}

// ---------------------------------------------------------------------------------------------------------
// Peer-memory form of the partitioned H.v: one arena per rank, exported with CUDA IPC and mapped by every other rank
// ---------------------------------------------------------------------------------------------------------
bool bh_dist_peer_wanted(const bh_ctx* ctx)
{
    static const int enabled = getenv("BH_DIST_PEER") ? atoi(getenv("BH_DIST_PEER")) : 1;
    return enabled && ctx->partitioned && ctx->world > 1 && ctx->world <= 8 && ctx->h_tab.chain && ctx->m >= 3 && !getenv("BH_DIST_ALLGATHER");
}

void bh_dist_arena_release(bh_ctx* ctx)
{
    if (!ctx->d_arena) return;
    cudaStreamSynchronize(ctx->stream);
    for (int p = 0; p < (int)ctx->peer_arena.size(); ++p)
        if (p != ctx->rank && ctx->peer_arena[p]) cudaIpcCloseMemHandle(ctx->peer_arena[p]);
    ctx->peer_arena.clear();
    ctx->peer_ready = false;
    for (cudaStream_t st : ctx->pull_stream)
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (cudaEvent_t e : ctx->pull_done)
        if (e) cudaEventDestroy(e);
    ctx->pull_stream.clear();
    ctx->pull_done.clear();
    // every rank has unmapped the others' arenas before any arena is freed
    if (ctx->nccl_comm && ctx->d_barrier) {
        g_nccl.AllReduce(ctx->d_barrier, ctx->d_barrier, 1, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(ctx->d_arena);
    ctx->d_arena = nullptr;
    ctx->arena_ncv = 0;
    ctx->d_V = ctx->d_w = ctx->d_f = nullptr;
    for (int q = 0; q < 3; ++q) ctx->d_cheb[q] = nullptr;
    ctx->ws_ncv = 0;
}

// Arena layout (doubles): V (ncv + 1) ld | w ld | f ld | cheb0 ld | cheb1 ld | cheb2 ld.  Collective: every rank calls it with
// the same ncv at the same point of the solve.  Growing keeps w, f and the Chebyshev buffers (not the basis, like the
// ordinary workspace).
int bh_dist_arena(bh_ctx* ctx, int ncv)
{
    if (ctx->d_arena && ncv <= ctx->arena_ncv) return BH_OK;
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = ctx->world;
    const int64_t ld = ctx->ld;
    const size_t ndoubles = (size_t)ld * (ncv + 1 + 5);
    double* fresh = nullptr;
    BH_CUDA(ctx, cudaMalloc(&fresh, sizeof(double) * ndoubles));
    BH_CUDA(ctx, cudaMemsetAsync(fresh, 0, sizeof(double) * ndoubles, ctx->stream));
    double* nw = fresh + (size_t)ld * (ncv + 1);
    if (ctx->d_arena) {  // keep the small vectors (pointer roles may have been swapped: copy by role)
        BH_CUDA(ctx, cudaMemcpyAsync(nw, ctx->d_w, sizeof(double) * ld, cudaMemcpyDeviceToDevice, ctx->stream));
        BH_CUDA(ctx, cudaMemcpyAsync(nw + ld, ctx->d_f, sizeof(double) * ld, cudaMemcpyDeviceToDevice, ctx->stream));
        for (int q = 0; q < 3; ++q)
            BH_CUDA(ctx, cudaMemcpyAsync(nw + (2 + q) * ld, ctx->d_cheb[q], sizeof(double) * ld, cudaMemcpyDeviceToDevice, ctx->stream));
        BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        bh_dist_arena_release(ctx);
    }
    if (!ctx->d_barrier) {
        BH_CUDA(ctx, cudaMalloc(&ctx->d_barrier, sizeof(double) * 2));
        BH_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, sizeof(double) * 2, ctx->stream));
    }
    ctx->d_arena = fresh;
    ctx->arena_ncv = ncv;
    ctx->d_V = fresh;
    ctx->d_w = nw;
    ctx->d_f = nw + ld;
    for (int q = 0; q < 3; ++q) ctx->d_cheb[q] = nw + (2 + q) * ld;
    ctx->ws_ncv = ncv;
    // export my arena, gather everybody's handle, map the others
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t mine;
    ctx->peer_ready = false;
    ctx->peer_arena.assign(W, nullptr);
    ctx->peer_arena[ctx->rank] = fresh;
    const bool exported = cudaIpcGetMemHandle(&mine, fresh) == cudaSuccess;
    if (!exported) {
        cudaGetLastError();
        std::memset(&mine, 0, sizeof(mine));
    }
    unsigned char* d_h = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_h, 64 * (size_t)W + 8));
    BH_CUDA(ctx, cudaMemcpyAsync(d_h + 64 * (size_t)ctx->rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    BH_NCCL(ctx, g_nccl.AllGather(d_h + 64 * (size_t)ctx->rank, d_h, 64, ncclUint8, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    std::vector<cudaIpcMemHandle_t> all(W);
    BH_CUDA(ctx, cudaMemcpyAsync(all.data(), d_h, 64 * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_h);
    bool ok = exported;
    for (int p = 0; p < W && ok; ++p) {
        if (p == ctx->rank) continue;
        bool zero = true;
        for (size_t b = 0; b < 64; ++b) zero = zero && reinterpret_cast<const unsigned char*>(&all[p])[b] == 0;
        void* ptr = nullptr;
        if (zero || cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            break;
        }
        ctx->peer_arena[p] = ptr;
    }
    // the form is used only if EVERY rank could map every arena (one all-reduce of a flag; also the first barrier)
    double flag = ok ? 0.0 : 1.0;
    BH_CUDA(ctx, cudaMemcpyAsync(ctx->d_barrier + 1, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BH_NCCL(ctx, g_nccl.AllReduce(ctx->d_barrier + 1, ctx->d_barrier + 1, 1, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    BH_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_barrier + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->peer_ready = (flag == 0.0);
    if (getenv("BH_DIST_VERBOSE"))
        fprintf(stderr, "[bh] rank %d arena %.1f MB (ncv %d): peer-memory H.v %s\n", ctx->rank, ndoubles * 8e-6, ncv,
                ctx->peer_ready ? "enabled" : "NOT available (CUDA IPC failed on some rank): halo exchange form");
    return BH_OK;
}

// Copy-engine pulls of this rank's halo ranges out of the peers' arenas (x lives at offset x_off of every arena) into d_xfull,
// on the communication stream, ordered after everything enqueued on the context's stream so far (the barrier included).
int bh_dist_pull_begin(bh_ctx* ctx, int64_t x_off)
{
    const int W = ctx->world;
    if ((int)ctx->pull_stream.size() < W) {
        int prio_lo = 0, prio_hi = 0;
        BH_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        ctx->pull_stream.resize(W, nullptr);
        ctx->pull_done.resize(W, nullptr);
        for (int p = 0; p < W; ++p) {
            if (p == ctx->rank) continue;
            BH_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->pull_stream[p], cudaStreamNonBlocking, prio_hi));
            BH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pull_done[p], cudaEventDisableTiming));
        }
    }
    BH_CUDA(ctx, cudaEventRecord(ctx->ev_x_ready, ctx->stream));
    std::vector<char> used(W, 0);
    for (const auto& r : ctx->halo_recv) {
        if (!used[r.peer]) {
            BH_CUDA(ctx, cudaStreamWaitEvent(ctx->pull_stream[r.peer], ctx->ev_x_ready, 0));
            used[r.peer] = 1;
        }
        const double* src = static_cast<const double*>(ctx->peer_arena[r.peer]) + x_off + (r.off - ctx->ld * r.peer);
        BH_CUDA(ctx, cudaMemcpyAsync(ctx->d_xfull + r.off, src, sizeof(double) * (size_t)r.count, cudaMemcpyDefault, ctx->pull_stream[r.peer]));
    }
    for (int p = 0; p < W; ++p)
        if (used[p]) BH_CUDA(ctx, cudaEventRecord(ctx->pull_done[p], ctx->pull_stream[p]));
    return BH_OK;
}

// the context's stream waits for the pulls of bh_dist_pull_begin
int bh_dist_pull_end(bh_ctx* ctx)
{
    std::vector<char> used(ctx->world, 0);
    for (const auto& r : ctx->halo_recv) used[r.peer] = 1;
    for (int p = 0; p < ctx->world; ++p)
        if (used[p]) BH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pull_done[p], 0));
    return BH_OK;
}

int bh_dist_barrier(bh_ctx* ctx)
{
    BH_NCCL(ctx, g_nccl.AllReduce(ctx->d_barrier, ctx->d_barrier, 1, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    return BH_OK;
}

extern "C" int bh_dist_finalize(bh_ctx* ctx)
{
    if (!ctx) return BH_ERR_ARG;
    bh_dist_arena_release(ctx);
    if (ctx->d_barrier) { cudaFree(ctx->d_barrier); ctx->d_barrier = nullptr; }
    bh_dist_release_halo(ctx);
    if (ctx->comm_stream) {
        cudaStreamSynchronize(ctx->comm_stream);
        if (ctx->nccl_comm2) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm2));
        ctx->nccl_comm2 = nullptr;
        cudaStreamDestroy(ctx->comm_stream);
        ctx->comm_stream = nullptr;
        if (ctx->ev_x_ready) cudaEventDestroy(ctx->ev_x_ready);
        if (ctx->ev_halo_done) cudaEventDestroy(ctx->ev_halo_done);
        ctx->ev_x_ready = ctx->ev_halo_done = nullptr;
    }
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


// ---------------------------------------------------------------------------------------------------------
// Overlapped halo exchange for the row-partitioned matrix-free H.v (chains).  The ncclAllGather of the whole vector
// (world * ld doubles into every rank, every H.v) is what made two GPUs slower than one; a rank's hops only read
// part of the other slices.  Once per bh_setup_partitioned every rank marks the 4096-row chunks its hops read
// (k_mark_halo_chain), the flag arrays are all-gathered, and each rank derives the ranges it must send to / receive
// from every peer.  Per H.v the ranges travel by grouped ncclSend / ncclRecv on a SECOND communicator and stream,
// straight from the caller's local vector into d_xfull at their global offsets, while the context's stream computes the
// hops whose source is local (k_hv_free_chain_part<.., 1>); the remote hops follow once the exchange has completed.
// ---------------------------------------------------------------------------------------------------------

int bh_dist_plan_halo(bh_ctx* ctx)
{
    bh_dist_release_halo(ctx);
    if (!ctx->partitioned || ctx->world < 2 || !ctx->h_tab.chain || ctx->m < 3) return BH_OK;
    if (getenv("BH_DIST_ALLGATHER")) return BH_OK;  // the north-star baseline: full ncclAllGather per H.v
    if (!g_nccl.CommSplit) return BH_OK;             // old NCCL: keep the all-gather path
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = ctx->world;
    const int64_t per = ctx->ld;  // slice length (equal on every rank, a multiple of 32)
    const int64_t nchunks = (per * W + HALO_CHUNK - 1) / HALO_CHUNK;
    if (!ctx->comm_stream) {
        ncclComm_t c2;
        BH_NCCL(ctx, g_nccl.CommSplit(static_cast<ncclComm_t>(ctx->nccl_comm), 0, ctx->rank, &c2, nullptr));
        ctx->nccl_comm2 = c2;
        int prio_lo = 0, prio_hi = 0;
        BH_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        BH_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi));
        BH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_x_ready, cudaEventDisableTiming));
        BH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_halo_done, cudaEventDisableTiming));
    }
    // flags of every rank: [W][nchunks] bytes
    unsigned char* d_flags = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_flags, (size_t)nchunks * W));
    BH_CUDA(ctx, cudaMemsetAsync(d_flags, 0, (size_t)nchunks * W, ctx->stream));
    BH_TRY(bh_mark_halo_chunks(ctx, d_flags + (size_t)ctx->rank * nchunks));
    BH_NCCL(ctx, g_nccl.AllGather(d_flags + (size_t)ctx->rank * nchunks, d_flags, (size_t)nchunks, ncclUint8,
                                  static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    std::vector<unsigned char> flags((size_t)nchunks * W);
    BH_D2H(ctx, flags.data(), d_flags, flags.size());
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_flags);
    // what this rank receives from / sends to every peer (halo_plan.h: host-only, tested on the CPU)
    {
        std::vector<BhHaloRange> recv, send;
        bh_halo_plan(flags.data(), nchunks, W, ctx->rank, per, ctx->D, 8, recv, send);
        for (const auto& r : recv) ctx->halo_recv.push_back({r.peer, r.off, r.count});
        for (const auto& r : send) ctx->halo_send.push_back({r.peer, r.off, r.count});
    }
    ctx->halo_recv_elems = 0;
    for (const auto& r : ctx->halo_recv) ctx->halo_recv_elems += r.count;
    BH_TRY(bh_build_remote_hops(ctx));
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->halo_ready = true;
    if (getenv("BH_DIST_VERBOSE"))
        fprintf(stderr, "[bh] rank %d of %d halo plan: %zu recv ranges (%.1f MB = %.3f D), %zu send ranges, %.2f remote hops per row\n", ctx->rank, W,
                ctx->halo_recv.size(), ctx->halo_recv_elems * 8e-6, (double)ctx->halo_recv_elems / (double)ctx->D, ctx->halo_send.size(),
                (double)ctx->rem_nnz / (double)std::max<int64_t>(ctx->nloc, 1));
    return BH_OK;
}

int bh_dist_halo_begin(bh_ctx* ctx, const double* x_local)
{
    // the exchange may start once x is final and the previous H.v has finished reading d_xfull: both are prior work of
    // the context's stream
    BH_CUDA(ctx, cudaEventRecord(ctx->ev_x_ready, ctx->stream));
    BH_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_x_ready, 0));
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm2);
    BH_NCCL(ctx, g_nccl.GroupStart());
    for (const auto& r : ctx->halo_send)
        BH_NCCL(ctx, g_nccl.Send(x_local + (r.off - ctx->row0), (size_t)r.count, ncclDouble, r.peer, comm, ctx->comm_stream));
    for (const auto& r : ctx->halo_recv)
        BH_NCCL(ctx, g_nccl.Recv(ctx->d_xfull + r.off, (size_t)r.count, ncclDouble, r.peer, comm, ctx->comm_stream));
    BH_NCCL(ctx, g_nccl.GroupEnd());
    BH_CUDA(ctx, cudaEventRecord(ctx->ev_halo_done, ctx->comm_stream));
    return BH_OK;
}

int bh_dist_halo_end(bh_ctx* ctx)
{
    BH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_halo_done, 0));
    return BH_OK;
}
