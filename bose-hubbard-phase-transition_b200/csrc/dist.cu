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
#include <string>

#include "bh_internal.h"

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
    ctx->halo_ready = false;
    ctx->halo_send.clear();
    ctx->halo_recv.clear();
    ctx->halo_piece_off.clear();
    ctx->halo_recv_elems = 0;
}

extern "C" int bh_dist_finalize(bh_ctx* ctx)
{
    if (!ctx) return BH_ERR_ARG;
    bh_dist_release_halo(ctx);
    if (ctx->comm_stream) {
        cudaStreamSynchronize(ctx->comm_stream);
        if (ctx->nccl_comm2) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm2));
        ctx->nccl_comm2 = nullptr;
        cudaStreamDestroy(ctx->comm_stream);
        ctx->comm_stream = nullptr;
        if (ctx->ev_x_ready) cudaEventDestroy(ctx->ev_x_ready);
        ctx->ev_x_ready = nullptr;
        for (cudaEvent_t e : ctx->ev_piece) cudaEventDestroy(e);
        ctx->ev_piece.clear();
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
// Pipelined halo exchange for the row-partitioned matrix-free H.v (chains).  The ncclAllGather of the whole vector
// (world * ld doubles into every rank, every H.v) is what made two GPUs slower than one in round 1: a rank's hops only read
// part of the other slices, and nothing was overlapped.  Here a rank sweeps its slice in `pieces` row ranges.  Once per
// bh_setup_partitioned every rank marks, per piece, the 4096-row chunks of the global vector that the hops of that piece
// read outside the slice (k_mark_halo_chain); the flag arrays are all-gathered and each rank derives, per piece, the ranges
// it must receive (chunks no earlier piece already fetched) and the ranges it must send to every peer.  Per H.v the pieces'
// ranges travel as `pieces` grouped ncclSend / ncclRecv batches on a SECOND communicator and a high-priority stream, from
// the caller's local vector straight into d_xfull at their global offsets; the context's stream waits for batch q only
// before sweeping piece q, so the exchange of the later pieces hides behind the sweep of the earlier ones.  Because a hop
// across bond (q, q+1) shifts the LEX rank monotonically, consecutive row pieces read (mostly) consecutive remote ranges:
// measured on the hop graph the new chunks per piece are roughly total / pieces (tools/halo_fraction.py).
// ---------------------------------------------------------------------------------------------------------
#define HALO_CHUNK 4096

int bh_dist_plan_halo(bh_ctx* ctx)
{
    bh_dist_release_halo(ctx);
    if (!ctx->partitioned || ctx->world < 2 || !ctx->h_tab.chain || ctx->m < 3) return BH_OK;
    if (getenv("BH_DIST_ALLGATHER")) return BH_OK;  // the north-star baseline: full ncclAllGather per H.v
    if (!g_nccl.CommSplit) return BH_OK;             // old NCCL: keep the all-gather path
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = ctx->world;
    const int NP = std::max(1, std::min(16, getenv("BH_HALO_PIECES") ? atoi(getenv("BH_HALO_PIECES")) : 4));
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
    }
    while ((int)ctx->ev_piece.size() < NP) {
        cudaEvent_t e;
        BH_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_piece.push_back(e);
    }
    // row pieces of this rank's slice (multiples of 256 rows)
    ctx->halo_piece_off.assign(NP + 1, ctx->nloc);
    for (int q = 0; q < NP; ++q) ctx->halo_piece_off[q] = std::min<int64_t>(ctx->nloc, (ctx->nloc * q / NP) / 256 * 256);
    // flags of every (rank, piece): [W][NP][nchunks] bytes
    const size_t per_rank = (size_t)NP * nchunks;
    unsigned char* d_flags = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_flags, per_rank * W));
    BH_CUDA(ctx, cudaMemsetAsync(d_flags, 0, per_rank * W, ctx->stream));
    for (int q = 0; q < NP; ++q)
        BH_TRY(bh_mark_halo_chunks(ctx, ctx->halo_piece_off[q], ctx->halo_piece_off[q + 1] - ctx->halo_piece_off[q],
                                   d_flags + (size_t)ctx->rank * per_rank + (size_t)q * nchunks));
    BH_NCCL(ctx, g_nccl.AllGather(d_flags + (size_t)ctx->rank * per_rank, d_flags, per_rank, ncclUint8,
                                  static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    std::vector<unsigned char> flags(per_rank * W);
    BH_D2H(ctx, flags.data(), d_flags, flags.size());
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_flags);
    // piece q of a reader fetches the chunks it reads that none of its earlier pieces fetched
    for (int r = 0; r < W; ++r)
        for (int64_t c = 0; c < nchunks; ++c) {
            bool seen = false;
            for (int q = 0; q < NP; ++q) {
                unsigned char& f = flags[(size_t)r * per_rank + (size_t)q * nchunks + c];
                if (f && seen) f = 0;
                seen = seen || f;
            }
        }
    // ranges of rank `owner`'s slice that piece q of rank `reader` fetches: runs of flagged chunks clipped to the owner's
    // slice; runs separated by fewer than 4 clean chunks are merged (fewer, larger messages)
    auto ranges = [&](int reader, int q, int owner, std::vector<bh_ctx::HaloRange>& out, int peer) {
        const int64_t lo = per * owner, hi = std::min<int64_t>(ctx->D, per * (owner + 1));
        if (hi <= lo) return;
        const unsigned char* f = flags.data() + (size_t)reader * per_rank + (size_t)q * nchunks;
        const int64_t c0 = lo / HALO_CHUNK, c1 = (hi + HALO_CHUNK - 1) / HALO_CHUNK;
        int64_t run0 = -1, last = -1;
        auto flush = [&]() {
            if (run0 < 0) return;
            const int64_t a = std::max(lo, run0 * HALO_CHUNK), b = std::min(hi, (last + 1) * HALO_CHUNK);
            if (b > a) out.push_back({peer, a, b - a});
            run0 = -1;
        };
        for (int64_t c = c0; c < c1; ++c) {
            if (!f[c]) continue;
            if (run0 >= 0 && c - last > 4) flush();
            if (run0 < 0) run0 = c;
            last = c;
        }
        flush();
    };
    ctx->halo_send.assign(NP, {});
    ctx->halo_recv.assign(NP, {});
    ctx->halo_recv_elems = 0;
    size_t nranges = 0;
    for (int q = 0; q < NP; ++q) {
        for (int p = 0; p < W; ++p) {
            if (p == ctx->rank) continue;
            ranges(ctx->rank, q, p, ctx->halo_recv[q], p);  // what my piece q reads from p's slice
            ranges(p, q, ctx->rank, ctx->halo_send[q], p);  // what p's piece q reads from mine
        }
        for (const auto& r : ctx->halo_recv[q]) ctx->halo_recv_elems += r.count;
        nranges += ctx->halo_recv[q].size();
    }
    ctx->halo_ready = true;
    if (getenv("BH_DIST_VERBOSE")) {
        std::string per_piece;
        for (int q = 0; q < NP; ++q) {
            int64_t e = 0;
            for (const auto& r : ctx->halo_recv[q]) e += r.count;
            per_piece += " " + std::to_string(e * 8 / 1000000) + "MB";
        }
        fprintf(stderr, "[bh] rank %d halo plan: %d pieces, %zu recv ranges, %.1f MB = %.3f D per H.v; per piece:%s\n", ctx->rank, NP, nranges,
                ctx->halo_recv_elems * 8e-6, (double)ctx->halo_recv_elems / (double)ctx->D, per_piece.c_str());
    }
    return BH_OK;
}

int bh_dist_halo_begin(bh_ctx* ctx, const double* x_local)
{
    // the exchange may start once x is final and the previous H.v has finished reading d_xfull: both are prior work of
    // the context's stream
    BH_CUDA(ctx, cudaEventRecord(ctx->ev_x_ready, ctx->stream));
    BH_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_x_ready, 0));
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm2);
    const int NP = (int)ctx->halo_recv.size();
    for (int q = 0; q < NP; ++q) {
        if (!ctx->halo_send[q].empty() || !ctx->halo_recv[q].empty()) {
            BH_NCCL(ctx, g_nccl.GroupStart());
            for (const auto& r : ctx->halo_send[q])
                BH_NCCL(ctx, g_nccl.Send(x_local + (r.off - ctx->row0), (size_t)r.count, ncclDouble, r.peer, comm, ctx->comm_stream));
            for (const auto& r : ctx->halo_recv[q])
                BH_NCCL(ctx, g_nccl.Recv(ctx->d_xfull + r.off, (size_t)r.count, ncclDouble, r.peer, comm, ctx->comm_stream));
            BH_NCCL(ctx, g_nccl.GroupEnd());
        }
        BH_CUDA(ctx, cudaEventRecord(ctx->ev_piece[q], ctx->comm_stream));
    }
    return BH_OK;
}

int bh_dist_halo_wait(bh_ctx* ctx, int piece)
{
    BH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_piece[piece], 0));
    return BH_OK;
}
