// hv.cu -- K3 (stored-matrix H.v) and K4 (matrix-free H.v).
//
// Replaces Spectra::SparseGenMatProd::perform_op -> Eigen's serial CSC scatter product
// (reference external/spectra/include/Spectra/MatOp/SparseGenMatProd.h:81-86,
//  include/Eigen/src/SparseCore/SparseDenseProduct.h:86-107).
//
// K3  streams a SELL-32-sigma copy of the materialised H once per product: a warp owns a slice of 32 rows, lane = row,
//     column-major entries (perfectly coalesced index / value loads), each lane sums its row in column order
//     (deterministic).  Algorithmic bytes: 12*nnz + 4*(D+1) + 16*D.  (The CSR-stream, TMA-ring, unrolled, hybrid and
//     split variants of round 1 were measured slower and removed in round 2: DESIGN.md section 10, git history.)
// K4  stores no matrix: thread k re-derives row k from the packed state (8 B) with the O(1) incremental
//     rank, so the HBM traffic is x, y, the packed states and the U-diagonal only.
#include <algorithm>
#include <cstdlib>

#include "device_utils.cuh"

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// K3, SELL-32 variant: one warp per slice of 32 rows, lane = row, entries column-major inside the slice.
// Index / value loads are perfectly coalesced streaming loads (evict-first), no shared memory, full
// occupancy; each lane accumulates its row in column order with FMAs.
// ---------------------------------------------------------------------------------------------
#ifndef SELL_BATCH
#define SELL_BATCH 12
#endif
#ifndef SELL_MINB
#define SELL_MINB 4
#endif
__global__ void __launch_bounds__(256, SELL_MINB)
k_hv_sell(int64_t D, int64_t nslices, const int* __restrict__ sptr, const int* __restrict__ srow,
          const int* __restrict__ scol, const double* __restrict__ sval, const double* __restrict__ x,
          double* __restrict__ y, BhEpilogue ep)
{
    const int lane = threadIdx.x & 31;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nslices; s += wstride) {
        const int base = __ldg(sptr + s), end = __ldg(sptr + s + 1);
        double acc = 0.0;
        for (int p = base + lane; p < end; p += 32 * SELL_BATCH) {
            int c[SELL_BATCH];
            double v[SELL_BATCH], xv[SELL_BATCH];
#pragma unroll
            for (int u = 0; u < SELL_BATCH; ++u) {
                const bool ok = p + 32 * u < end;
                c[u] = ok ? __ldcs(scol + p + 32 * u) : -1;
                v[u] = ok ? __ldcs(sval + p + 32 * u) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < SELL_BATCH; ++u) xv[u] = (c[u] >= 0) ? __ldg(x + c[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < SELL_BATCH; ++u) acc = fma(v[u], xv[u], acc);
        }
        const int r = __ldg(srow + (s << 5) + lane);  // sigma-sorted slot -> row
        if (r >= 0) {
            double out = ep.s1 * acc;
            if (ep.s2 != 0.0) out = fma(ep.s2, x[r], out);
            if (ep.z) out = fma(ep.s3, ep.z[r], out);
            y[r] = out;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------------------------
typedef void (*hv_free_fn_t)(const BhTables*, int64_t, int64_t, const uint64_t*, const double*, double, double, double,
                             const double*, double*, BhEpilogue);

// K4, bond-list variant: the per-thread rank prefixes live in shared memory ([site][thread], conflict-free),
// so the hop loop runs over the lattice's actual bond list (uniform across the block) instead of all site
// pairs, and one kernel serves every m.  ~4x fewer instructions per row than the unrolled variant.
#define HVF_THREADS 256
__global__ void __launch_bounds__(HVF_THREADS)
k_hv_free_bonds(const BhTables* __restrict__ gtab, int64_t row0, int64_t D, const uint64_t* __restrict__ states,
                const double* __restrict__ dU, double cJ, double cU, double cmu, const double* __restrict__ x,
                double* __restrict__ y, BhEpilogue ep)
{
    __shared__ BhTables t;
    __shared__ int sdn[BH_MAX_SITES][HVF_THREADS];
    __shared__ int sup[BH_MAX_SITES][HVF_THREADS];
    bh_stage_tables(&t, gtab);
    const int tid = threadIdx.x;
    const int m = t.m, nb = t.nbonds;
    const double shift = __dmul_rn(-(double)t.n, cmu);
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + tid; l < D; l += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = row0 + l;
        const uint64_t s = states[l];
        int R = t.n, adn = 0, aup = 0;
        for (int q = 0; q < m; ++q) {
            sdn[q][tid] = adn;
            sup[q][tid] = aup;
            R -= bh_occ(s, q);
            if (q < m - 1) {
                const int f0 = t.f[q][R];
                adn += (R >= 1 ? t.f[q][R - 1] : f0) - f0;
                aup += t.f[q][R + 1] - f0;
            }
        }
        double acc = 0.0;
#pragma unroll 4
        for (int b = 0; b < nb; ++b) {
            const int bd = t.bond[b];
            const int dst = bd & 15, src = (bd >> 4) & 15, w = bd >> 8;
            const int ns = bh_occ(s, src);
            const int delta = (dst < src) ? sdn[src][tid] - sdn[dst][tid] : sup[dst][tid] - sup[src][tid];
            const double xv = ns ? __ldg(x + (int)k + delta) : 0.0;
            acc = fma((double)w * t.sq[(bh_occ(s, dst) + 1) * ns], xv, acc);
        }
        const double diag = __dadd_rn(__dmul_rn(dU[l], cU), shift);
        const double xk = x[k];
        double out = ep.s1 * (diag * xk - cJ * acc);
        if (ep.s2 != 0.0) out = fma(ep.s2, xk, out);
        if (ep.z) out = fma(ep.s3, ep.z[l], out);
        y[l] = out;
    }
}

// K4, chain-specialised variant (the reference CLI only ever builds the closed chain, src/analysis.cpp:219-220):
// a nearest-neighbour hop across bond (q, q+1) changes the rank by ONE table difference that depends on
// (q, R_q) only, so the row is produced in a single sweep over the sites with no prefix arrays; the periodic
// bond uses the accumulated totals.  ~3x fewer instructions per row than the bond-list kernel.
#ifndef CHAIN_MIN_BLOCKS
#define CHAIN_MIN_BLOCKS 4
#endif
template <int M, bool CLOSED>
__global__ void __launch_bounds__(256, CHAIN_MIN_BLOCKS)
k_hv_free_chain(const BhTables* __restrict__ gtab, int64_t row0, int64_t D, const uint64_t* __restrict__ states,
                const double* __restrict__ dU, double cJ, double cU, double cmu, const double* __restrict__ x,
                double* __restrict__ y, BhEpilogue ep)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const double shift = __dmul_rn(-(double)t.n, cmu);
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < D; l += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = row0 + l;
        const uint64_t s = states[l];
        const int kk = (int)k;
        const int n0 = bh_occ(s, 0);
        int R = t.n - n0, nprev = n0, tdn = 0, tup = 0;
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < M - 1; ++q) {
            const int nnext = bh_occ(s, q + 1);
            const int2 gh = t.gh[q][R];  // .x: boson moves q+1 -> q, .y: q -> q+1
            const double xa = nnext ? __ldg(x + (kk + gh.x)) : 0.0;
            const double xb = nprev ? __ldg(x + (kk + gh.y)) : 0.0;
            acc = fma(t.sq[(nprev + 1) * nnext], xa, acc);
            acc = fma(t.sq[(nnext + 1) * nprev], xb, acc);
            tdn += gh.x;
            tup += gh.y;
            R -= nnext;
            nprev = nnext;
        }
        if (CLOSED) {
            const int nl = nprev;  // occupation of the last site
            const double xa = nl ? __ldg(x + (kk + tdn)) : 0.0;  // M-1 -> 0
            const double xb = n0 ? __ldg(x + (kk + tup)) : 0.0;  // 0 -> M-1
            acc = fma(t.sq[(n0 + 1) * nl], xa, acc);
            acc = fma(t.sq[(nl + 1) * n0], xb, acc);
        }
        const double diag = __dadd_rn(__dmul_rn(dU[l], cU), shift);
        const double xv = x[k];
        double out = ep.s1 * (diag * xv - (2.0 * cJ) * acc);
        if (ep.s2 != 0.0) out = fma(ep.s2, xv, out);
        if (ep.z) out = fma(ep.s3, ep.z[l], out);
        y[l] = out;
    }
}

// Row-partitioned form of the chain kernel (one large eigensolve over several GPUs, BASELINE.json config 5): the hops of a
// row are split by where their source element lives.  This kernel takes the hops whose source is in this rank's own slice
// [row0, row0 + D) -- x is then the LOCAL slice, addressed through a pointer shifted by -row0 -- plus the diagonal and the
// epilogue, and runs while the halo exchange is in flight: the single sweep of k_hv_free_chain with a range test per hop.
// The hops whose source lies in another rank's slice are a stored CSR matrix applied afterwards (k_hv_remote).
template <int M, bool CLOSED>
__global__ void __launch_bounds__(256, CHAIN_MIN_BLOCKS)
k_hv_free_chain_part(const BhTables* __restrict__ gtab, int64_t row0, int64_t D, const uint64_t* __restrict__ states,
                     const double* __restrict__ dU, double cJ, double cU, double cmu, const double* __restrict__ x,
                     double* __restrict__ y, BhEpilogue ep)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const double shift = __dmul_rn(-(double)t.n, cmu);
    const unsigned lo = (unsigned)row0, len = (unsigned)D;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < D; l += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = row0 + l;
        const uint64_t s = states[l];
        const int kk = (int)k;
        const int n0 = bh_occ(s, 0);
        int R = t.n - n0, nprev = n0, tdn = 0, tup = 0;
        double acc = 0.0;
        auto hop = [&](int cond, int tgt, double amp) {
            const bool local = ((unsigned)tgt - lo) < len;
            const bool take = cond && local;
            const double xv = take ? __ldg(x + tgt) : 0.0;
            acc = fma(amp, xv, acc);
        };
#pragma unroll
        for (int q = 0; q < M - 1; ++q) {
            const int nnext = bh_occ(s, q + 1);
            const int2 gh = t.gh[q][R];
            hop(nnext, kk + gh.x, t.sq[(nprev + 1) * nnext]);
            hop(nprev, kk + gh.y, t.sq[(nnext + 1) * nprev]);
            tdn += gh.x;
            tup += gh.y;
            R -= nnext;
            nprev = nnext;
        }
        if (CLOSED) {
            const int nl = nprev;
            hop(nl, kk + tdn, t.sq[(n0 + 1) * nl]);
            hop(n0, kk + tup, t.sq[(nl + 1) * n0]);
        }
        const double diag = __dadd_rn(__dmul_rn(dU[l], cU), shift);
        const double xv = x[k];
        double out = ep.s1 * (diag * xv - (2.0 * cJ) * acc);
        if (ep.s2 != 0.0) out = fma(ep.s2, xv, out);
        if (ep.z) out = fma(ep.s3, ep.z[l], out);
        y[l] = out;
    }
}

// Which 4096-row chunks of the global vector hold a source element of some hop of this rank's rows: flags[chunk] = 1
// (the halo plan of dist.cu is built from these flags once per bh_setup_partitioned).
#define HALO_CHUNK_SHIFT 12
template <int M, bool CLOSED>
__global__ void __launch_bounds__(256)
k_mark_halo_chain(const BhTables* __restrict__ gtab, int64_t row0, int64_t D, const uint64_t* __restrict__ states,
                  unsigned char* __restrict__ flags)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const unsigned lo = (unsigned)row0, len = (unsigned)D;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < D; l += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t s = states[l];
        const int kk = (int)(row0 + l);
        const int n0 = bh_occ(s, 0);
        int R = t.n - n0, nprev = n0, tdn = 0, tup = 0;
        auto mark = [&](int cond, int tgt) {
            if (cond && !(((unsigned)tgt - lo) < len)) flags[tgt >> HALO_CHUNK_SHIFT] = 1;
        };
#pragma unroll
        for (int q = 0; q < M - 1; ++q) {
            const int nnext = bh_occ(s, q + 1);
            const int2 gh = t.gh[q][R];
            mark(nnext, kk + gh.x);
            mark(nprev, kk + gh.y);
            tdn += gh.x;
            tup += gh.y;
            R -= nnext;
            nprev = nnext;
        }
        if (CLOSED) {
            mark(nprev, kk + tdn);
            mark(n0, kk + tup);
        }
    }
}

template <bool CLOSED>
static hv_free_fn_t hv_chain_kernel(int m)
{
    switch (m) {
        case 3: return k_hv_free_chain<3, CLOSED>;
        case 4: return k_hv_free_chain<4, CLOSED>;
        case 5: return k_hv_free_chain<5, CLOSED>;
        case 6: return k_hv_free_chain<6, CLOSED>;
        case 7: return k_hv_free_chain<7, CLOSED>;
        case 8: return k_hv_free_chain<8, CLOSED>;
        case 9: return k_hv_free_chain<9, CLOSED>;
        case 10: return k_hv_free_chain<10, CLOSED>;
        case 11: return k_hv_free_chain<11, CLOSED>;
        case 12: return k_hv_free_chain<12, CLOSED>;
        case 13: return k_hv_free_chain<13, CLOSED>;
        case 14: return k_hv_free_chain<14, CLOSED>;
        case 15: return k_hv_free_chain<15, CLOSED>;
        case 16: return k_hv_free_chain<16, CLOSED>;
    }
    return nullptr;
}


template <bool CLOSED>
static hv_free_fn_t hv_chain_part_kernel(int m)
{
    switch (m) {
        case 3: return k_hv_free_chain_part<3, CLOSED>;
        case 4: return k_hv_free_chain_part<4, CLOSED>;
        case 5: return k_hv_free_chain_part<5, CLOSED>;
        case 6: return k_hv_free_chain_part<6, CLOSED>;
        case 7: return k_hv_free_chain_part<7, CLOSED>;
        case 8: return k_hv_free_chain_part<8, CLOSED>;
        case 9: return k_hv_free_chain_part<9, CLOSED>;
        case 10: return k_hv_free_chain_part<10, CLOSED>;
        case 11: return k_hv_free_chain_part<11, CLOSED>;
        case 12: return k_hv_free_chain_part<12, CLOSED>;
        case 13: return k_hv_free_chain_part<13, CLOSED>;
        case 14: return k_hv_free_chain_part<14, CLOSED>;
        case 15: return k_hv_free_chain_part<15, CLOSED>;
        case 16: return k_hv_free_chain_part<16, CLOSED>;
    }
    return nullptr;
}

typedef void (*mark_fn_t)(const BhTables*, int64_t, int64_t, const uint64_t*, unsigned char*);
template <bool CLOSED>
static mark_fn_t mark_halo_kernel(int m)
{
    switch (m) {
        case 3: return k_mark_halo_chain<3, CLOSED>;
        case 4: return k_mark_halo_chain<4, CLOSED>;
        case 5: return k_mark_halo_chain<5, CLOSED>;
        case 6: return k_mark_halo_chain<6, CLOSED>;
        case 7: return k_mark_halo_chain<7, CLOSED>;
        case 8: return k_mark_halo_chain<8, CLOSED>;
        case 9: return k_mark_halo_chain<9, CLOSED>;
        case 10: return k_mark_halo_chain<10, CLOSED>;
        case 11: return k_mark_halo_chain<11, CLOSED>;
        case 12: return k_mark_halo_chain<12, CLOSED>;
        case 13: return k_mark_halo_chain<13, CLOSED>;
        case 14: return k_mark_halo_chain<14, CLOSED>;
        case 15: return k_mark_halo_chain<15, CLOSED>;
        case 16: return k_mark_halo_chain<16, CLOSED>;
    }
    return nullptr;
}

// The remote hops of every local row as a CSR matrix (built once per bh_setup_partitioned): MODE 0 counts them per row,
// MODE 1 writes (source rank, amplitude) at the scanned offsets, in the order of the sweep.
template <int M, bool CLOSED, int MODE>
__global__ void __launch_bounds__(256)
k_remote_hops(const BhTables* __restrict__ gtab, int64_t row0, int64_t D, const uint64_t* __restrict__ states,
              int* __restrict__ ptr_or_count, int* __restrict__ col, double* __restrict__ amp)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const unsigned lo = (unsigned)row0, len = (unsigned)D;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < D; l += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t s = states[l];
        const int kk = (int)(row0 + l);
        const int n0 = bh_occ(s, 0);
        int R = t.n - n0, nprev = n0, tdn = 0, tup = 0;
        int pos = (MODE == 1) ? ptr_or_count[l] : 0;
        auto hop = [&](int cond, int tgt, int code) {
            if (cond && !(((unsigned)tgt - lo) < len)) {
                if (MODE == 1) {
                    col[pos] = tgt;
                    amp[pos] = t.sq[code];
                }
                ++pos;
            }
        };
#pragma unroll
        for (int q = 0; q < M - 1; ++q) {
            const int nnext = bh_occ(s, q + 1);
            const int2 gh = t.gh[q][R];
            hop(nnext, kk + gh.x, (nprev + 1) * nnext);
            hop(nprev, kk + gh.y, (nnext + 1) * nprev);
            tdn += gh.x;
            tup += gh.y;
            R -= nnext;
            nprev = nnext;
        }
        if (CLOSED) {
            const int nl = nprev;
            hop(nl, kk + tdn, (n0 + 1) * nl);
            hop(n0, kk + tup, (nl + 1) * n0);
        }
        if (MODE == 0) ptr_or_count[l] = pos;
    }
}

// y[l] += coef * sum_e amp[e] x[col[e]] over the remote hops of row l (x = the exchanged full-length buffer)
__global__ void __launch_bounds__(256)
k_hv_remote(int64_t D, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ amp,
            const double* __restrict__ x, double* __restrict__ y, double coef)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= D) return;
    const int a = __ldg(ptr + l), b = __ldg(ptr + l + 1);
    if (a == b) return;
    double acc = 0.0;
    for (int e = a; e < b; ++e) acc = fma(__ldcs(amp + e), __ldg(x + __ldcs(col + e)), acc);
    y[l] = fma(coef, acc, y[l]);
}

typedef void (*remote_fn_t)(const BhTables*, int64_t, int64_t, const uint64_t*, int*, int*, double*);
template <bool CLOSED, int MODE>
static remote_fn_t remote_hops_kernel(int m)
{
    switch (m) {
        case 3: return k_remote_hops<3, CLOSED, MODE>;
        case 4: return k_remote_hops<4, CLOSED, MODE>;
        case 5: return k_remote_hops<5, CLOSED, MODE>;
        case 6: return k_remote_hops<6, CLOSED, MODE>;
        case 7: return k_remote_hops<7, CLOSED, MODE>;
        case 8: return k_remote_hops<8, CLOSED, MODE>;
        case 9: return k_remote_hops<9, CLOSED, MODE>;
        case 10: return k_remote_hops<10, CLOSED, MODE>;
        case 11: return k_remote_hops<11, CLOSED, MODE>;
        case 12: return k_remote_hops<12, CLOSED, MODE>;
        case 13: return k_remote_hops<13, CLOSED, MODE>;
        case 14: return k_remote_hops<14, CLOSED, MODE>;
        case 15: return k_remote_hops<15, CLOSED, MODE>;
        case 16: return k_remote_hops<16, CLOSED, MODE>;
    }
    return nullptr;
}

int bh_build_remote_hops(bh_ctx* ctx)
{
    if (ctx->d_rem_ptr) { cudaFree(ctx->d_rem_ptr); ctx->d_rem_ptr = nullptr; }
    if (ctx->d_rem_col) { cudaFree(ctx->d_rem_col); ctx->d_rem_col = nullptr; }
    if (ctx->d_rem_amp) { cudaFree(ctx->d_rem_amp); ctx->d_rem_amp = nullptr; }
    ctx->rem_nnz = 0;
    const int64_t nloc = ctx->nloc;
    const bool closed = ctx->h_tab.chain == 2;
    BH_CUDA(ctx, cudaMalloc(&ctx->d_rem_ptr, sizeof(int) * (size_t)(nloc + 1)));
    BH_CUDA(ctx, cudaMemsetAsync(ctx->d_rem_ptr, 0, sizeof(int) * (size_t)(nloc + 1), ctx->stream));
    if (nloc == 0) return BH_OK;
    int* d_cnt = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_cnt, sizeof(int) * (size_t)nloc));
    const int grid = (int)std::min<int64_t>(nblocks(nloc, 256), (int64_t)ctx->sm_count * 8);
    remote_fn_t f0 = closed ? remote_hops_kernel<true, 0>(ctx->m) : remote_hops_kernel<false, 0>(ctx->m);
    f0<<<grid, 256, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, nloc, ctx->d_states, d_cnt, nullptr, nullptr);
    BH_LAUNCHED(ctx);
    int64_t total = 0;
    BH_TRY(bh_exclusive_scan(ctx, nloc, d_cnt, ctx->d_rem_ptr, &total));
    cudaFree(d_cnt);
    if (total >= ((int64_t)1 << 31)) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "remote hop matrix: more than 2^31 entries");
    ctx->rem_nnz = total;
    BH_CUDA(ctx, cudaMalloc(&ctx->d_rem_col, sizeof(int) * (size_t)std::max<int64_t>(total, 1)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_rem_amp, sizeof(double) * (size_t)std::max<int64_t>(total, 1)));
    remote_fn_t f1 = closed ? remote_hops_kernel<true, 1>(ctx->m) : remote_hops_kernel<false, 1>(ctx->m);
    f1<<<grid, 256, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, nloc, ctx->d_states, ctx->d_rem_ptr, ctx->d_rem_col, ctx->d_rem_amp);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}

// flags_dev[(world * ld) >> 12 chunks] <- 1 where this rank's rows read a remote chunk (chains only); see dist.cu
int bh_mark_halo_chunks(bh_ctx* ctx, unsigned char* flags_dev)
{
    if (!ctx->h_tab.chain || ctx->m < 3) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "halo plan: chains only");
    if (ctx->nloc == 0) return BH_OK;
    mark_fn_t fn = (ctx->h_tab.chain == 2) ? mark_halo_kernel<true>(ctx->m) : mark_halo_kernel<false>(ctx->m);
    const int grid = (int)std::min<int64_t>(nblocks(ctx->nloc, 256), (int64_t)ctx->sm_count * 8);
    fn<<<grid, 256, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, ctx->nloc, ctx->d_states, flags_dev);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}

// Gershgorin bounds of the spectrum: row k gives diag_k -/+ |cJ| sum_hops w sqrt((n_dst+1) n_src).  Used to scale
// the Chebyshev filter of the accelerated solver (the upper bound must be rigorous: the filter explodes above it).
__global__ void __launch_bounds__(256)
k_gersh(const BhTables* __restrict__ gtab, int64_t nloc, const uint64_t* __restrict__ states, const double* __restrict__ dU,
        double cJ, double cU, double cmu, double* __restrict__ part /* [2][gridDim.x] */)
{
    __shared__ BhTables t;
    __shared__ double shi[8], slo[8];
    bh_stage_tables(&t, gtab);
    double hi = -1e300, lo = 1e300;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < nloc; l += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t s = states[l];
        double off = 0.0;
        for (int b = 0; b < t.nbonds; ++b) {
            const int bd = t.bond[b];
            const int dst = bd & 15, src = (bd >> 4) & 15, w = bd >> 8;
            off += (double)w * t.sq[(bh_occ(s, dst) + 1) * bh_occ(s, src)];
        }
        off *= fabs(cJ);
        const double diag = dU[l] * cU - (double)t.n * cmu;
        hi = fmax(hi, diag + off);
        lo = fmin(lo, diag - off);
    }
    for (int o = 16; o > 0; o >>= 1) {
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    }
    if ((threadIdx.x & 31) == 0) { shi[threadIdx.x >> 5] = hi; slo[threadIdx.x >> 5] = lo; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) { hi = fmax(hi, shi[q]); lo = fmin(lo, slo[q]); }
        part[blockIdx.x] = hi;
        part[gridDim.x + blockIdx.x] = -lo;  // stored negated so that one max-reduction serves both
    }
}

__global__ void k_gersh_final(int nb, const double* __restrict__ part, double* __restrict__ out2)
{
    double a = -1e300, b = -1e300;
    for (int i = threadIdx.x; i < nb; i += 32) { a = fmax(a, part[i]); b = fmax(b, part[nb + i]); }
    for (int o = 16; o > 0; o >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (threadIdx.x == 0) { out2[0] = a; out2[1] = b; }
}

int bh_spectrum_bounds(bh_ctx* ctx, double cJ, double cU, double cmu, double* lo, double* hi)
{
    if (ctx->user_matrix) return bh_fail(ctx, BH_ERR_STATE, "spectrum bounds need a model context");
    BH_TRY(bh_ensure_workspace(ctx, 0));
    const int nb = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks(ctx->nloc, 256), (int64_t)ctx->sm_count * 4));
    double* part = ctx->d_part;
    k_gersh<<<nb, 256, 0, ctx->stream>>>(ctx->d_tab, ctx->nloc, ctx->d_states, ctx->d_dU, cJ, cU, cmu, part);
    k_gersh_final<<<1, 32, 0, ctx->stream>>>(nb, part, part + 2 * (size_t)nb);
    ctx->launches += 2;
    BH_TRY(bh_dist_allreduce_max(ctx, part + 2 * (size_t)nb, 2));
    double h2[2];
    BH_D2H(ctx, h2, part + 2 * (size_t)nb, sizeof(double) * 2);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *hi = h2[0];
    *lo = -h2[1];
    return BH_OK;
}

// generic epilogue for the kernels that do not fuse it: y = s1*y + s2*x + s3*z
__global__ void k_epilogue(int64_t n, double* __restrict__ y, const double* __restrict__ x, BhEpilogue ep)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double out = ep.s1 * y[r] + ep.s2 * x[r];
        if (ep.z) out = fma(ep.s3, ep.z[r], out);
        y[r] = out;
    }
}

int bh_launch_hv(bh_ctx* ctx, double cJ, double cU, double cmu, int kernel_in, const double* x, double* y, const BhEpilogue& ep)
{
    int kernel = kernel_in;
    bool fused = false;
    const bool plain = (ep.s1 == 1.0 && ep.s2 == 0.0 && ep.z == nullptr);
    const int64_t D = ctx->D;
    // algorithmic bytes of SURVEY.md section 8d (stored: 12 nnz + 4 (D + 1) + 16 D; matrix-free: 16 D per local row range)
    const bool prof_stored = (kernel_in == BH_HV_STORED || kernel_in == BH_HV_USER);
    BhProfScope prof(ctx, prof_stored ? BH_PROF_HV_STORED : BH_PROF_HV_FREE,
                     prof_stored ? 12.0 * (double)ctx->nnzH + 4.0 * (double)(D + 1) + 16.0 * (double)D : 16.0 * (double)ctx->nloc);
    // BH_HV_HYBRID was a round-1 experiment (stored slices + matrix-free rows in one launch, not faster); the id is kept in
    // the ABI and selects the matrix-free kernel
    if (kernel == BH_HV_HYBRID) kernel = BH_HV_MATRIX_FREE;
    if (ctx->partitioned && kernel != BH_HV_MATRIX_FREE)
        return bh_fail(ctx, BH_ERR_STATE, "a row-partitioned context has no stored matrix: use BH_HV_MATRIX_FREE");
    if (ctx->user_matrix != (kernel == BH_HV_USER))
        return bh_fail(ctx, BH_ERR_STATE, "kernel BH_HV_USER needs bh_load_matrix, the other kernels need bh_setup");
    if (kernel == BH_HV_USER || kernel == BH_HV_STORED) {
        if (kernel == BH_HV_STORED) BH_TRY(bh_materialise_sell(ctx, cJ, cU, cmu));
        const int64_t ns = ctx->sell_nslices;
        const int grid = (int)std::min<int64_t>((ns + 7) / 8, (int64_t)ctx->sm_count * 8);
        k_hv_sell<<<grid, 256, 0, ctx->stream>>>(D, ns, ctx->d_sell_ptr, ctx->d_sell_row, ctx->d_sell_col, ctx->d_sell_valH, x, y, ep);
        fused = true;
        BH_LAUNCHED(ctx);
    } else if (kernel == BH_HV_MATRIX_FREE) {
        // row-partitioned context: x is the local slice; exchange it (NCCL all-gather) and read the full vector
        const double* xin = x;
        const int64_t nloc = ctx->nloc;
        if (ctx->partitioned && ctx->peer_ready && ctx->halo_ready && ctx->d_arena && x >= ctx->d_arena &&
            x + ctx->ld <= ctx->d_arena + (size_t)ctx->ld * (ctx->arena_ncv + 1 + 5)) {
            // Peer-memory form (default for chains): every rank keeps its vectors at the same offsets of an arena the others have
            // mapped with CUDA IPC.  After a barrier (all ranks have finished the kernels that wrote their part of x) the copy
            // engines PULL the ranges of the halo plan straight out of the owners' arenas over NVLink -- no NCCL kernel, no SM
            // taken from the sweep -- while the SMs compute the hops whose source is local; the stored remote hops follow.
            // (Loading the remote elements inside the sweep itself was measured slower: NVLink latency stalls a kernel that is
            // already latency-bound, 0.36 s per solve against 0.29 s)
            BH_TRY(bh_dist_barrier(ctx));
            BH_TRY(bh_dist_pull_begin(ctx, x - ctx->d_arena));
            if (nloc > 0) {
                const bool closed = ctx->h_tab.chain == 2;
                hv_free_fn_t f1 = closed ? hv_chain_part_kernel<true>(ctx->m) : hv_chain_part_kernel<false>(ctx->m);
                const int grid = (int)std::min<int64_t>(nblocks(nloc, 256), (int64_t)ctx->sm_count * 8);
                f1<<<grid, 256, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, nloc, ctx->d_states, ctx->d_dU, cJ, cU, cmu, x - ctx->row0, y, ep);
                BH_LAUNCHED(ctx);
            }
            BH_TRY(bh_dist_pull_end(ctx));
            if (nloc > 0 && ctx->rem_nnz > 0) {
                k_hv_remote<<<nblocks(nloc, 256), 256, 0, ctx->stream>>>(nloc, ctx->d_rem_ptr, ctx->d_rem_col, ctx->d_rem_amp, ctx->d_xfull, y,
                                                                         ep.s1 * (-2.0 * cJ));
                BH_LAUNCHED(ctx);
            }
            BH_CUDA(ctx, cudaGetLastError());
            return BH_OK;
        }
        if (ctx->partitioned && ctx->halo_ready && ctx->h_tab.chain && ctx->m >= 3) {
            // overlapped form: halo exchange on the communication stream, own-slice hops meanwhile, remote hops afterwards
            static const int ablate = getenv("BH_HALO_ABLATE") ? atoi(getenv("BH_HALO_ABLATE")) : 0;  // timing probes only
            if (!(ablate & 1)) BH_TRY(bh_dist_halo_begin(ctx, x));
            const bool closed = ctx->h_tab.chain == 2;
            // one CTA per 256 rows (not a persistent grid): the exchange kernel on the high-priority communication stream gets
            // SM slots as soon as the first CTAs retire, so that the two really overlap.  Measured at m = n = 14 on 2 GPUs (per
            // H.v): a persistent grid of 8 CTAs per SM serialises the exchange behind the sweep (596 us), 7 per SM leaves no room
            // for an NCCL CTA either (~420 us), one CTA per 256 rows 386 us = sweep 268 (with the 202 us exchange hidden) + remote
            // hops 92 + 26 exposed.  A pipelined variant (row pieces waiting only for their own remote chunks, one ordinary sweep
            // per piece) was built and measured slower: 4 grouped ncclSend/Recv batches take 334 us against 202 us for one.
            static const int per_sm = getenv("BH_HALO_GRID") ? atoi(getenv("BH_HALO_GRID")) : 0;  // > 0: persistent CTAs per SM
            const int grid = per_sm > 0 ? (int)std::min<int64_t>(nblocks(nloc, 256), (int64_t)ctx->sm_count * per_sm)
                                        : (int)std::max<int64_t>(1, nblocks(nloc, 256));
            if (nloc > 0 && !(ablate & 2)) {
                hv_free_fn_t f1 = closed ? hv_chain_part_kernel<true>(ctx->m) : hv_chain_part_kernel<false>(ctx->m);
                f1<<<grid, 256, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, nloc, ctx->d_states, ctx->d_dU, cJ, cU, cmu, x - ctx->row0, y, ep);
                BH_LAUNCHED(ctx);
            }
            if (!(ablate & 1)) BH_TRY(bh_dist_halo_end(ctx));
            if (nloc > 0 && !(ablate & 4) && ctx->rem_nnz > 0) {
                k_hv_remote<<<nblocks(nloc, 256), 256, 0, ctx->stream>>>(nloc, ctx->d_rem_ptr, ctx->d_rem_col, ctx->d_rem_amp, ctx->d_xfull, y,
                                                           ep.s1 * (-2.0 * cJ));
                BH_LAUNCHED(ctx);
            }
            BH_CUDA(ctx, cudaGetLastError());
            return BH_OK;
        }
        if (ctx->partitioned) {
            BH_TRY(bh_dist_allgather(ctx, x, ctx->d_xfull, ctx->ld));
            xin = ctx->d_xfull;
        }
        if (nloc > 0) {
            if (ctx->h_tab.chain) {  // nearest-neighbour chain: single sweep over the sites
                hv_free_fn_t fn = (ctx->h_tab.chain == 2) ? hv_chain_kernel<true>(ctx->m) : hv_chain_kernel<false>(ctx->m);
                int grid = (int)std::min<int64_t>(nblocks(nloc, 256), (int64_t)ctx->sm_count * 8);
                fn<<<grid, 256, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, nloc, ctx->d_states, ctx->d_dU, cJ, cU, cmu, xin, y, ep);
            } else {  // any lattice: bond list
                int grid = (int)std::min<int64_t>(nblocks(nloc, HVF_THREADS), (int64_t)ctx->sm_count * 5);
                k_hv_free_bonds<<<grid, HVF_THREADS, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, nloc, ctx->d_states, ctx->d_dU, cJ, cU, cmu, xin, y, ep);
            }
            fused = true;
            BH_LAUNCHED(ctx);
        }
    } else {
        return bh_fail(ctx, BH_ERR_ARG, "unknown H.v kernel");
    }
    if (!fused && !plain && ctx->nloc > 0) {
        // x here must be the LOCAL slice for the s2 term: the unfused kernels are never used on partitioned contexts
        k_epilogue<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->nloc, y, x, ep);
        BH_LAUNCHED(ctx);
    }
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}

extern "C" int bh_hv_dev(bh_ctx* ctx, double cJ, double cU, double cmu, int kernel, const double* x_dev, double* y_dev)
{
    if (!ctx || !ctx->D) return bh_fail(ctx, BH_ERR_STATE, "bh_hv_dev: call bh_setup first");
    if (!x_dev || !y_dev || x_dev == y_dev) return bh_fail(ctx, BH_ERR_ARG, "bh_hv_dev: bad vectors");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    return bh_launch_hv(ctx, cJ, cU, cmu, kernel, x_dev, y_dev);
}

extern "C" int bh_hv(bh_ctx* ctx, double cJ, double cU, double cmu, int kernel, int order, const double* x, double* y)
{
    if (!ctx || !ctx->D) return bh_fail(ctx, BH_ERR_STATE, "bh_hv: call bh_setup first");
    if (!x || !y || order < 0 || order > 2) return bh_fail(ctx, BH_ERR_ARG, "bh_hv: bad argument");
    if (ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_hv: use bh_hv_dev with the local slice on a row-partitioned context");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    BH_TRY(bh_ensure_staging(ctx));
    BH_TRY(bh_ensure_workspace(ctx, 0));
    const size_t bytes = sizeof(double) * ctx->D;
    BH_H2D(ctx, ctx->d_y, x, bytes);
    BH_TRY(bh_permute_vec(ctx, order, false, ctx->d_y, ctx->d_x));   // x_lex
    BH_TRY(bh_launch_hv(ctx, cJ, cU, cmu, kernel, ctx->d_x, ctx->d_w));
    BH_TRY(bh_permute_vec(ctx, order, true, ctx->d_w, ctx->d_y));    // y in `order`
    BH_D2H(ctx, y, ctx->d_y, bytes);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BH_OK;
}

extern "C" int bh_host_register(void* ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return BH_ERR_ARG;
    return cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? BH_OK : BH_ERR_CUDA;
}

extern "C" int bh_host_unregister(void* ptr)
{
    if (!ptr) return BH_ERR_ARG;
    return cudaHostUnregister(ptr) == cudaSuccess ? BH_OK : BH_ERR_CUDA;
}

extern "C" int bh_hv_algorithmic_bytes(bh_ctx* ctx, int kernel, int64_t* bytes)
{
    if (!ctx || !ctx->D || !bytes) return bh_fail(ctx, BH_ERR_STATE, "bh_hv_algorithmic_bytes: call bh_setup first");
    if (kernel == BH_HV_STORED && !ctx->user_matrix) BH_TRY(bh_build_hamiltonian(ctx));
    if (kernel == BH_HV_STORED || kernel == BH_HV_USER)
        *bytes = 12 * ctx->nnzH + 4 * (ctx->D + 1) + 16 * ctx->D;
    else
        *bytes = 16 * ctx->D;
    return BH_OK;
}

// ---------------------------------------------------------------------------------------------
// Spectra's LCG start vector (Util/SimpleRandom.h:30-64): seed_{i+1} = 16807 seed_i mod (2^31 - 1), seed_0 = 1,
// element i = seed_{i+1} / (2^31 - 1) - 0.5.  Each thread jumps ahead by modular exponentiation.
// ---------------------------------------------------------------------------------------------
#define LCG_CHUNK 64
__global__ void k_lcg_fill(int64_t first, int64_t n, double* __restrict__ out)
{
    const unsigned long long M = 2147483647ull;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t start = c * LCG_CHUNK;
    if (start >= n) return;
    unsigned long long seed = 1, base = 16807ull;
    for (unsigned long long p = (unsigned long long)(first + start); p; p >>= 1) {
        if (p & 1) seed = seed * base % M;
        base = base * base % M;
    }
    const int64_t end = min(start + LCG_CHUNK, n);
    for (int64_t i = start; i < end; ++i) {
        seed = seed * 16807ull % M;
        out[i] = (double)(long long)seed / (double)2147483647.0 - 0.5;
    }
}

extern "C" int bh_lcg_fill_dev(bh_ctx* ctx, double* x_dev, int64_t count)
{
    if (!ctx || !x_dev || count < 0) return bh_fail(ctx, BH_ERR_ARG, "bh_lcg_fill_dev: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return BH_OK;
    // a row-partitioned context fills its slice of the one global sequence
    const int64_t first = ctx->partitioned ? ctx->row0 : 0;
    k_lcg_fill<<<nblocks((count + LCG_CHUNK - 1) / LCG_CHUNK, 128), 128, 0, ctx->stream>>>(first, count, x_dev);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}
