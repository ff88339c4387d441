// batch.cu -- several grid points of the sweep share their H.v launches.
//
// The reference's sweep loop (src/analysis.cpp:302-343) solves every grid point independently; the points differ only
// in the coefficients (cJ, cU, cmu) of the same three operators.  bh_points(.., batch = B) runs B eigensolves in
// lockstep: each solve is the unmodified single-point code path (bh_point -> bh_lanczos) on a child context with its own
// Krylov workspace, executed by its own host thread, but only ONE thread runs at a time (a baton), all of them launch on
// the parent's stream, and a solve that reaches its Chebyshev filter parks there.  When every live solve is parked,
// the last arrival applies all the filters with ONE kernel per degree: the d intermediate vectors T_k of the B points
// are stored interleaved ([row][point]), so a hop computes its rank and amplitude once and fetches B consecutive
// doubles with one 16/32-byte access.  The per-row H.v kernel is bound by L1 tag look-ups of scattered gathers
// (12 cache lines per warp instruction, profiles/), which the interleaving amortises over the batch.
// Per point the arithmetic is the same sequence of operations as the single-point kernel: results are bit-identical.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "bh_internal.h"
#include "device_utils.cuh"
#include "lockstep_sched.h"

#define BH_MAX_BATCH 4

struct BatchReq {
    const double* x = nullptr;
    double* y = nullptr;
    double c = 0, e = 0, cJ = 0, cU = 0, cmu = 0;
    int d = 0;
};

struct bh_batch_hub {
    LockstepSched sched;  // who runs, who is parked (lockstep_sched.h)
    BatchReq req[BH_MAX_BATCH];
    double* d_il[3] = {nullptr, nullptr, nullptr};  // interleaved Chebyshev buffers, ld * BH_MAX_BATCH doubles each
    int64_t il_len = 0;
    int64_t batched_filters = 0, single_filters = 0;
};

// ---------------------------------------------------------------------------------------------
// Batched chain kernel: k_hv_free_chain (hv.cu) for NB vectors at once.
// ---------------------------------------------------------------------------------------------
struct BatchVecs {
    const double* xs[BH_MAX_BATCH];  // separate sources (first filter step) ...
    const double* xi;                // ... or the interleaved source
    double* ys[BH_MAX_BATCH];        // separate destinations (last filter step) ...
    double* yi;                      // ... or the interleaved destination
    const double* zs[BH_MAX_BATCH];  // T_{k-2}: separate (second step), interleaved, or none
    const double* zi;
    double twoJ[BH_MAX_BATCH], cU[BH_MAX_BATCH], shift[BH_MAX_BATCH];
    double s1[BH_MAX_BATCH], s2[BH_MAX_BATCH], s3[BH_MAX_BATCH];
};

// four interleaved doubles with ONE 256-bit access (sm_100: LDG.E.256 / STG.E.256): one L1 tag look-up per lane and hop
__device__ __forceinline__ void load_il(const double* __restrict__ base, int idx, double (&v)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(base + (size_t)idx * 4));
}

__device__ __forceinline__ void store_il(double* base, size_t row, const double (&v)[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(base + row * 4), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

__device__ __forceinline__ void store_il(double* base, size_t row, const double (&v)[2])
{
    *reinterpret_cast<double2*>(base + row * 2) = make_double2(v[0], v[1]);
}

template <int NB>
__device__ __forceinline__ void load_il(const double* __restrict__ base, int idx, double (&v)[NB])
{
    const double2* p = reinterpret_cast<const double2*>(base + (size_t)idx * NB);
#pragma unroll
    for (int h = 0; h < NB / 2; ++h) {
        const double2 t = __ldg(p + h);
        v[2 * h] = t.x;
        v[2 * h + 1] = t.y;
    }
}

template <int M, int NB, bool SRC_IL, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_hv_chain_batch(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states,
                 const double* __restrict__ dU, const __grid_constant__ BatchVecs a)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < D; l += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t s = states[l];
        const int kk = (int)l;
        const int n0 = bh_occ(s, 0);
        int R = t.n - n0, nprev = n0, tdn = 0, tup = 0;
        double acc[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = 0.0;
        auto hop = [&](int cond, int idx, double amp) {
            double v[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) v[b] = 0.0;
            if (cond) {
                if (SRC_IL) {
                    load_il(a.xi, idx, v);
                } else {
#pragma unroll
                    for (int b = 0; b < NB; ++b) v[b] = __ldg(a.xs[b] + idx);
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b] = fma(amp, v[b], acc[b]);
        };
#pragma unroll
        for (int q = 0; q < M - 1; ++q) {
            const int nnext = bh_occ(s, q + 1);
            const int2 gh = t.gh[q][R];  // .x: boson moves q+1 -> q, .y: q -> q+1
            hop(nnext, kk + gh.x, t.sq[(nprev + 1) * nnext]);
            hop(nprev, kk + gh.y, t.sq[(nnext + 1) * nprev]);
            tdn += gh.x;
            tup += gh.y;
            R -= nnext;
            nprev = nnext;
        }
        {
            const int nl = nprev;  // occupation of the last site; periodic bond
            hop(nl, kk + tdn, t.sq[(n0 + 1) * nl]);  // M-1 -> 0
            hop(n0, kk + tup, t.sq[(nl + 1) * n0]);  // 0 -> M-1
        }
        double xv[NB], zv[NB], out[NB];
        if (SRC_IL) {
            load_il(a.xi, kk, xv);
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) xv[b] = a.xs[b][l];
        }
        const double du = dU[l];
        const bool zil = a.zi != nullptr, zsep = a.zs[0] != nullptr;
        if (zil) load_il(a.zi, kk, zv);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const double diag = __dadd_rn(__dmul_rn(du, a.cU[b]), a.shift[b]);
            double o = a.s1[b] * (diag * xv[b] - a.twoJ[b] * acc[b]);
            if (a.s2[b] != 0.0) o = fma(a.s2[b], xv[b], o);
            if (zil) o = fma(a.s3[b], zv[b], o);
            else if (zsep) o = fma(a.s3[b], a.zs[b][l], o);
            out[b] = o;
        }
        if (a.yi) {
            store_il(a.yi, (size_t)l, out);
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) a.ys[b][l] = out[b];
        }
    }
}

typedef void (*batch_fn)(const BhTables*, int64_t, const uint64_t*, const double*, const BatchVecs);

template <int NB, bool SRC_IL, int MINB>
static batch_fn batch_kernel_m(int m)
{
    switch (m) {
        case 6: return k_hv_chain_batch<6, NB, SRC_IL, MINB>;
        case 7: return k_hv_chain_batch<7, NB, SRC_IL, MINB>;
        case 8: return k_hv_chain_batch<8, NB, SRC_IL, MINB>;
        case 9: return k_hv_chain_batch<9, NB, SRC_IL, MINB>;
        case 10: return k_hv_chain_batch<10, NB, SRC_IL, MINB>;
        case 11: return k_hv_chain_batch<11, NB, SRC_IL, MINB>;
        case 12: return k_hv_chain_batch<12, NB, SRC_IL, MINB>;
        case 13: return k_hv_chain_batch<13, NB, SRC_IL, MINB>;
        case 14: return k_hv_chain_batch<14, NB, SRC_IL, MINB>;
        case 15: return k_hv_chain_batch<15, NB, SRC_IL, MINB>;
        case 16: return k_hv_chain_batch<16, NB, SRC_IL, MINB>;
    }
    return nullptr;
}

static batch_fn batch_kernel(int m, int nb, bool src_il)
{
    static const int minb4 = getenv("BH_BATCH4_MINB") ? atoi(getenv("BH_BATCH4_MINB")) : 3;
    if (nb == 2) return src_il ? batch_kernel_m<2, true, 4>(m) : batch_kernel_m<2, false, 4>(m);
    if (nb == 4 && minb4 == 4) return src_il ? batch_kernel_m<4, true, 4>(m) : batch_kernel_m<4, false, 4>(m);
    if (nb == 4) return src_il ? batch_kernel_m<4, true, 3>(m) : batch_kernel_m<4, false, 3>(m);
    return nullptr;
}

bool bh_batch_supported(const bh_ctx* ctx, int kernel)
{
    return kernel == BH_HV_MATRIX_FREE && !ctx->user_matrix && !ctx->partitioned && ctx->h_tab.chain == 2 && ctx->m >= 6 &&
           ctx->m <= 16 && ctx->cheb_degree > 1 && !ctx->parent;
}

// The Chebyshev filters of the parked fibers fib[0..nb), nb in {2, 4}, applied together (see bh_lanczos, stage 2).
static int apply_filters(bh_ctx* parent, bh_batch_hub* hub, const int* fib, int nb)
{
    const int m = parent->m;
    const int64_t D = parent->D;
    const int d = hub->req[fib[0]].d;
    const int grid = (int)std::min<int64_t>((D + 255) / 256, (int64_t)parent->sm_count * 8);
    for (int k = 1; k <= d; ++k) {
        BatchVecs a;
        std::memset(&a, 0, sizeof(a));
        const bool last = (k == d);
        for (int b = 0; b < nb; ++b) {
            const BatchReq& r = hub->req[fib[b]];
            double s1, s2, s3 = 0.0;
            if (k == 1) {
                s1 = 1.0 / r.e; s2 = -r.c / r.e;
            } else {
                s1 = 2.0 / r.e; s2 = -2.0 * r.c / r.e; s3 = -1.0;
            }
            if (last && (d % 2 == 0)) { s1 = -s1; s2 = -s2; s3 = -s3; }
            a.s1[b] = s1; a.s2[b] = s2; a.s3[b] = s3;
            a.twoJ[b] = 2.0 * r.cJ;
            a.cU[b] = r.cU;
            a.shift[b] = -(double)parent->n * r.cmu;
            if (k == 1) a.xs[b] = r.x;
            if (k == 2) a.zs[b] = r.x;
            if (last) a.ys[b] = r.y;
        }
        if (k >= 2) a.xi = hub->d_il[(k - 1) % 3];
        if (k >= 3) a.zi = hub->d_il[(k - 2) % 3];
        if (!last) a.yi = hub->d_il[k % 3];
        batch_fn fn = batch_kernel(m, nb, k >= 2);
        if (!fn) return bh_fail(parent, BH_ERR_UNSUPPORTED, "batched H.v: unsupported chain length");
        {
            BhProfScope prof(parent, nb == 4 ? BH_PROF_HV_BATCH4 : BH_PROF_HV_BATCH2, 16.0 * (double)D * nb);
            fn<<<grid, 256, 0, parent->stream>>>(parent->d_tab, D, parent->d_states, parent->d_dU, a);
        }
        BH_LAUNCHED(parent);
    }
    BH_CUDA(parent, cudaGetLastError());
    hub->batched_filters += nb;
    return BH_OK;
}

// ---- what a launch does for a group of parked solves with the same degree ----
// One parked fiber alone: its filter through the ordinary single-vector kernel on its own buffers.
static int apply_single(bh_ctx* parent, bh_batch_hub* hub, int f)
{
    const BatchReq& q = hub->req[f];
    bh_ctx* owner = parent->children[f];
    const double* tkm2 = q.x;
    const double* tkm1 = nullptr;
    int rc = BH_OK;
    for (int k = 1; k <= q.d && rc == BH_OK; ++k) {
        BhEpilogue ep;
        if (k == 1) {
            ep.s1 = 1.0 / q.e; ep.s2 = -q.c / q.e;
        } else {
            ep.s1 = 2.0 / q.e; ep.s2 = -2.0 * q.c / q.e; ep.s3 = -1.0; ep.z = tkm2;
        }
        const bool last = (k == q.d);
        if (last && (q.d % 2 == 0)) { ep.s1 = -ep.s1; ep.s2 = -ep.s2; ep.s3 = -ep.s3; }
        double* dst = last ? q.y : owner->d_cheb[k % 3];
        const double* src = (k == 1) ? q.x : tkm1;
        rc = bh_launch_hv(owner, q.cJ, q.cU, q.cmu, BH_HV_MATRIX_FREE, src, dst, ep);
        if (k >= 2) tkm2 = tkm1;
        tkm1 = dst;
    }
    if (rc != BH_OK) parent->err = owner->err;
    hub->single_filters++;
    return rc;
}

// Launch callback of the scheduler: the parked solves grp[0..ng) asked for the same degree; quadruples and pairs share
// their launches, a leftover runs alone.  Called with the scheduler lock held, by the thread that holds the baton.
static int launch_group(bh_ctx* parent, bh_batch_hub* hub, const int* grp, int ng)
{
    int pos = 0, rc = BH_OK;
    while (ng - pos >= 2 && rc == BH_OK) {
        const int nb = (ng - pos >= 4) ? 4 : 2;
        rc = apply_filters(parent, hub, grp + pos, nb);
        pos += nb;
    }
    if (rc == BH_OK && pos < ng) rc = apply_single(parent, hub, grp[pos]);
    return rc;
}

// Called by fiber `me` (holding the baton) from bh_lanczos: apply T_d((H - c) / e) to x.  *handled = false means "run the
// ordinary path yourself" (the other solves have finished); otherwise the launches of this fiber's filter have been
// enqueued on the shared stream (by this thread or by another one) when the call returns.
int bh_batch_filter(bh_ctx* child, const double* x, double* y, double c, double e, double cJ, double cU, double cmu, int d,
                    bool* handled)
{
    bh_ctx* parent = child->parent;
    bh_batch_hub* hub = parent->hub;
    const int me = child->fiber;
    // the request record is read by the launching thread only after this fiber has parked (under the scheduler lock)
    BatchReq& r = hub->req[me];
    r.x = x; r.y = y; r.c = c; r.e = e; r.cJ = cJ; r.cU = cU; r.cmu = cmu; r.d = d;
    int status = BH_OK;
    *handled = hub->sched.request(me, d, [&](const int* grp, int ng) { return launch_group(parent, hub, grp, ng); }, &status);
    if (status != BH_OK && !parent->err.empty()) child->err = parent->err;  // the shared launches failed: keep their message
    if (!*handled) hub->single_filters++;
    return status;
}

void bh_batch_release(bh_ctx* ctx)
{
    for (bh_ctx* c : ctx->children) {
        if (!c) continue;
        bh_release_workspace(c);
        delete c;
    }
    ctx->children.clear();
    if (ctx->hub) {
        for (int q = 0; q < 3; ++q)
            if (ctx->hub->d_il[q]) cudaFree(ctx->hub->d_il[q]);
        delete ctx->hub;
        ctx->hub = nullptr;
    }
}

static int ensure_children(bh_ctx* ctx, int nb)
{
    if (!ctx->hub) ctx->hub = new bh_batch_hub();
    bh_batch_hub* hub = ctx->hub;
    const int64_t need = ctx->ld * BH_MAX_BATCH;
    if (hub->il_len < need) {
        for (int q = 0; q < 3; ++q) {
            if (hub->d_il[q]) cudaFree(hub->d_il[q]);
            hub->d_il[q] = nullptr;
            BH_CUDA(ctx, cudaMalloc(&hub->d_il[q], sizeof(double) * need));
            BH_CUDA(ctx, cudaMemsetAsync(hub->d_il[q], 0, sizeof(double) * need, ctx->stream));
        }
        hub->il_len = need;
    }
    while ((int)ctx->children.size() < nb) {
        bh_ctx* c = new bh_ctx(*ctx);  // aliases the model arrays (states, dU, tables); owns only its workspace
        c->parent = ctx;
        c->children.clear();
        c->hub = nullptr;
        c->own_stream = false;
        c->fiber = (int)ctx->children.size();
        c->launches = c->h2d_bytes = c->d2h_bytes = 0;
        c->small_ws = nullptr;
        c->prof_on = false;  // event records live on the parent (BhProfScope)
        c->prof.clear();
        c->prof_free.clear();
        c->d_V = c->d_w = c->d_f = c->d_scal = c->d_part = c->d_small = c->d_spdm_scratch = c->d_x = c->d_y = nullptr;
        c->d_counter = nullptr;
        c->h_pinned = nullptr;
        c->h_pinned_bytes = 0;
        c->spdm_scratch_bytes = 0;
        c->ws_ncv = 0;
        for (int q = 0; q < 3; ++q) c->d_cheb[q] = nullptr;
        c->d_hv_block = nullptr;
        c->hv_block_cols = 0;
        c->d_gram_part = nullptr;
        ctx->children.push_back(c);
    }
    return BH_OK;
}

// npoints grid points on nb (2..4) lockstep solves; a solve that finishes its point takes the next one from the list, so
// the batch stays full until the list runs out.  Results as bh_point, point by point.
int bh_points_lockstep(bh_ctx* ctx, int nb, int64_t npoints, const double* cJ, const double* cU, const double* cmu, int nb_eigen,
                       int kernel, double* out3, bh_eigs_info* infos)
{
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const auto t_begin = std::chrono::steady_clock::now();
    BH_TRY(ensure_children(ctx, nb));
    const auto t_children = std::chrono::steady_clock::now();
    bh_batch_hub* hub = ctx->hub;
    hub->sched.reset(nb);
    int rcs[BH_MAX_BATCH] = {BH_OK, BH_OK, BH_OK, BH_OK};
    int64_t next_point = 0;  // guarded by the baton (only the running fiber touches it)
    bool stop = false;
    std::vector<std::thread> threads;
    for (int i = 0; i < nb; ++i) {
        bh_ctx* c = ctx->children[i];
        c->stream = ctx->stream;
        c->cheb_degree = ctx->cheb_degree;
        threads.emplace_back([=, &rcs, &next_point, &stop] {
            hub->sched.wait_first_turn(i);
            for (;;) {
                if (stop || next_point >= npoints) break;
                const int64_t p = next_point++;
                int rc;
                try {
                    rc = bh_point(c, cJ[p], cU[p], cmu[p], nb_eigen, kernel, out3 + 3 * p, nullptr, nullptr, infos ? infos + p : nullptr);
                } catch (...) {
                    rc = bh_fail(c, BH_ERR_STATE, "exception inside a lockstep solve");
                }
                if (rc != BH_OK) {
                    rcs[i] = rc;
                    stop = true;
                }
            }
            // leave: pass the baton; if every remaining live solve is parked, their filters are launched first
            hub->sched.finish(i, [&](const int* grp, int ng) { return launch_group(ctx, hub, grp, ng); });
        });
    }
    for (auto& t : threads) t.join();
    int rc = BH_OK;
    for (int i = 0; i < nb; ++i) {
        bh_ctx* c = ctx->children[i];
        ctx->launches += c->launches;
        ctx->h2d_bytes += c->h2d_bytes;
        ctx->d2h_bytes += c->d2h_bytes;
        c->launches = c->h2d_bytes = c->d2h_bytes = 0;
        if (rcs[i] != BH_OK && rc == BH_OK) {
            rc = rcs[i];
            ctx->err = c->err;
        }
    }
    if (getenv("BH_BATCH_VERBOSE")) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "[bh] lockstep: %lld points, filters applied batched %lld / single %lld (cumulative); children %.3f s, solves %.3f s\n",
                (long long)npoints, (long long)hub->batched_filters, (long long)hub->single_filters,
                std::chrono::duration<double>(t_children - t_begin).count(), std::chrono::duration<double>(t_end - t_children).count());
    }
    return rc;
}
