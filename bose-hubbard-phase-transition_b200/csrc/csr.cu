// csr.cu -- K2: GPU builder of the Hamiltonian's sparse pattern and values, and its exports.
//
// Replaces BH::fill_hopping / fill_interaction / fill_chemical / fixed_bosons_hamiltonian
// (reference src/hamiltonian.cpp:170-256) and the per-point sum H = H_fixed + H1*p1 + H2*p2
// (src/analysis.cpp:311).  Instead of tag + binary search per hop and a triplet sort, every row is
// produced independently: count -> scan -> fill, the column of each hop coming from the O(1) incremental
// rank (device_utils.cuh).  The stored pattern is JH u diagonal in LEX order with ascending columns --
// the constant pattern of the swept H (explicit diagonal kept).
#include <algorithm>

#include "device_utils.cuh"

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------------------------
// exclusive scan of int32 lengths -> int32 offsets (total checked against 2^31 on the host)
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8

__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

__global__ void k_scan_blocks(int64_t n, const int* __restrict__ in, int* __restrict__ out, long long* __restrict__ bsum)
{
    __shared__ int wsum[SCAN_THREADS / 32];
    const int64_t base = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        tsum += v[i];
    }
    const int incl = warp_incl_scan(tsum);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < SCAN_THREADS / 32) ? wsum[lane] : 0;
        w = warp_incl_scan(w);
        if (lane < SCAN_THREADS / 32) wsum[lane] = w;
    }
    __syncthreads();
    int excl = incl - tsum + (wid > 0 ? wsum[wid - 1] : 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == SCAN_THREADS - 1) bsum[blockIdx.x] = wsum[SCAN_THREADS / 32 - 1];
}

__global__ void k_scan_sums(int nb, long long* __restrict__ bsum, long long* __restrict__ total)
{
    // one warp walks the block sums with a running carry (nb is a few thousand at most)
    long long carry = 0;
    for (int base = 0; base < nb; base += 32) {
        const int i = base + threadIdx.x;
        long long v = (i < nb) ? bsum[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += u;
        }
        if (i < nb) bsum[i] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_scan_add(int64_t n, int* __restrict__ out, const long long* __restrict__ bsum,
                           const long long* __restrict__ total)
{
    const int64_t base = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    const int off = (int)bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) out[base + i] += off;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (int)(*total);
}

// out[0..n] = exclusive prefix sums of in[0..n); returns the total
int bh_exclusive_scan(bh_ctx* ctx, int64_t n, const int* d_in, int* d_out, int64_t* total)
{
    const int nb = nblocks(n, SCAN_THREADS * SCAN_ITEMS);
    long long* d_bsum = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_bsum, sizeof(long long) * (nb + 1)));
    k_scan_blocks<<<nb, SCAN_THREADS, 0, ctx->stream>>>(n, d_in, d_out, d_bsum);
    k_scan_sums<<<1, 32, 0, ctx->stream>>>(nb, d_bsum, d_bsum + nb);
    k_scan_add<<<nb, SCAN_THREADS, 0, ctx->stream>>>(n, d_out, d_bsum, d_bsum + nb);
    ctx->launches += 3;
    long long t = 0;
    BH_D2H(ctx, &t, d_bsum + nb, sizeof(long long));
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bsum);
    *total = t;
    return BH_OK;
}

// ---------------------------------------------------------------------------------------------
// K2 count / fill
// ---------------------------------------------------------------------------------------------
__global__ void k_row_count(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states,
                            int* __restrict__ rowlen)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D) return;
    const uint64_t s = states[k];
    int c = 1;  // the diagonal slot
    for (int src = 0; src < t.m; ++src) {
        if (bh_occ(s, src) == 0) continue;
        for (int dst = 0; dst < t.m; ++dst) c += (t.w[dst][src] != 0);
    }
    rowlen[k] = c;
}

#define BH_MAX_ROW 242

__global__ void k_row_fill(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states,
                           const int* __restrict__ rowptr, int* __restrict__ col, double* __restrict__ valJ,
                           int* __restrict__ diagpos)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D) return;
    const uint64_t s = states[k];
    // rank prefix arrays (dynamic m: local memory, set-up only)
    int dn[BH_MAX_SITES], up[BH_MAX_SITES];
    {
        int R = t.n, adn = 0, aup = 0;
        for (int q = 0; q < t.m; ++q) {
            dn[q] = adn;
            up[q] = aup;
            R -= bh_occ(s, q);
            if (q < t.m - 1) {
                const int f0 = t.f[q][R];
                adn += (R >= 1 ? t.f[q][R - 1] : f0) - f0;
                aup += t.f[q][R + 1] - f0;
            }
        }
    }
    int c[BH_MAX_ROW];
    double v[BH_MAX_ROW];
    int len = 0;
    c[len] = (int)k;
    v[len] = 0.0;
    ++len;
    for (int src = 0; src < t.m; ++src) {
        const int ns = bh_occ(s, src);
        if (ns == 0) continue;
        for (int dst = 0; dst < t.m; ++dst) {
            const int w = t.w[dst][src];
            if (!w) continue;
            const int tgt = (int)k + (dst < src ? dn[src] - dn[dst] : up[dst] - up[src]);
            // reference: -J * sqrt((n_dst + 1) * n_src), pushed w times and summed (J = 1 here)
            const double a = -(t.sq[(bh_occ(s, dst) + 1) * ns]);
            double acc = a;
            for (int r = 1; r < w; ++r) acc = __dadd_rn(acc, a);
            // insertion sort by column
            int p = len++;
            while (p > 0 && c[p - 1] > tgt) {
                c[p] = c[p - 1];
                v[p] = v[p - 1];
                --p;
            }
            c[p] = tgt;
            v[p] = acc;
        }
    }
    const int base = rowptr[k];
    for (int i = 0; i < len; ++i) {
        col[base + i] = c[i];
        valJ[base + i] = v[i];
        if (c[i] == (int)k) diagpos[k] = base + i;
    }
}

int bh_build_hamiltonian(bh_ctx* ctx)
{
    if (ctx->d_rowptr) return BH_OK;  // built lazily, once (the matrix-free paths never need it)
    if (ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "a row-partitioned context has no stored matrix");
    const int64_t D = ctx->D;
    int* d_len = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_len, sizeof(int) * D));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_rowptr, sizeof(int) * (D + 1 + 64)));  // slack: tiles copy whole 16-byte groups
    BH_CUDA(ctx, cudaMemsetAsync(ctx->d_rowptr, 0, sizeof(int) * (D + 1 + 64), ctx->stream));
    k_row_count<<<nblocks(D, 256), 256, 0, ctx->stream>>>(ctx->d_tab, D, ctx->d_states, d_len);
    BH_LAUNCHED(ctx);
    int64_t total = 0;
    BH_TRY(bh_exclusive_scan(ctx, D, d_len, ctx->d_rowptr, &total));
    cudaFree(d_len);
    if (total >= ((int64_t)1 << 31))
        return bh_fail(ctx, BH_ERR_UNSUPPORTED, "stored Hamiltonian has >= 2^31 entries; use the matrix-free kernel");
    ctx->nnzH = total;
    ctx->nnzJ = total - D;
    // +8 elements of slack: the H.v kernel may read whole aligned groups past the end
    BH_CUDA(ctx, cudaMalloc(&ctx->d_col, sizeof(int) * (total + 8)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_valJ, sizeof(double) * (total + 8)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_valH, sizeof(double) * (total + 8)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_diagpos, sizeof(int) * D));
    BH_CUDA(ctx, cudaMemsetAsync(ctx->d_col + total, 0, sizeof(int) * 8, ctx->stream));
    BH_CUDA(ctx, cudaMemsetAsync(ctx->d_valJ + total, 0, sizeof(double) * 8, ctx->stream));
    BH_CUDA(ctx, cudaMemsetAsync(ctx->d_valH + total, 0, sizeof(double) * 8, ctx->stream));
    k_row_fill<<<nblocks(D, 128), 128, 0, ctx->stream>>>(ctx->d_tab, D, ctx->d_states, ctx->d_rowptr, ctx->d_col,
                                                         ctx->d_valJ, ctx->d_diagpos);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    ctx->valH_valid = false;
    return BH_OK;
}

// ---------------------------------------------------------------------------------------------
// per-point materialisation  H = JH*cJ + UH*cU + uH*cmu  on the fixed pattern (src/analysis.cpp:311)
// ---------------------------------------------------------------------------------------------
__global__ void k_materialise(int64_t D, int n, const int* __restrict__ rowptr, const int* __restrict__ diagpos,
                              const double* __restrict__ valJ, const double* __restrict__ dU, double cJ, double cU,
                              double cmu, double* __restrict__ valH)
{
    // one warp per 32 rows; lanes stride over the contiguous entry range of those rows
    const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (row0 >= D) return;
    const int lane = threadIdx.x & 31;
    const int64_t rend = min(row0 + 32, D);
    const int e0 = rowptr[row0], e1 = rowptr[rend];
    for (int e = e0 + lane; e < e1; e += 32) valH[e] = __dmul_rn(valJ[e], cJ);
    __syncwarp();
    const int64_t r = row0 + lane;
    if (r < rend) {
        const int p = diagpos[r];
        // Eigen evaluates (JHcJ + UH*cU) + uH*cmu entry-wise; uH holds -n (mu = 1)
        const double a = __dadd_rn(__dmul_rn(valJ[p], cJ), __dmul_rn(dU[r], cU));
        valH[p] = __dadd_rn(a, __dmul_rn(-(double)n, cmu));
    }
}

int bh_materialise_H(bh_ctx* ctx, double cJ, double cU, double cmu)
{
    BH_TRY(bh_build_hamiltonian(ctx));
    if (ctx->valH_valid && ctx->cur_cJ == cJ && ctx->cur_cU == cU && ctx->cur_cmu == cmu) return BH_OK;
    const int wpb = 8;
    const int64_t nwarps = (ctx->D + 31) / 32;
    k_materialise<<<nblocks(nwarps, wpb), wpb * 32, 0, ctx->stream>>>(ctx->D, ctx->n, ctx->d_rowptr, ctx->d_diagpos,
                                                                        ctx->d_valJ, ctx->d_dU, cJ, cU, cmu, ctx->d_valH);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    ctx->cur_cJ = cJ;
    ctx->cur_cU = cU;
    ctx->cur_cmu = cmu;
    ctx->valH_valid = true;
    return BH_OK;
}

// ---------------------------------------------------------------------------------------------
// SELL-32 copy of the stored matrix (sliced ELLPACK, C = 32, sigma = 1: rows keep their LEX order so that the
// j-th entries of the 32 rows of a slice are mostly the same hop applied to neighbouring states and the
// gathers of x fall into a few cache lines).
// ---------------------------------------------------------------------------------------------
// sigma-sorting: inside every window of SIGMA consecutive rows the rows are ordered by decreasing length (stable),
// so that the 32 rows of a slice have nearly equal lengths (padding 16.8 % -> 6.0 % at sigma = 128 on the m=n=12
// chain) while staying close enough in LEX order for the x gathers to share cache lines.  slot -> row map, -1 = none.
__global__ void k_sell_perm(int64_t D, int sigma, const int* __restrict__ rowptr, int* __restrict__ slot_row)
{
    extern __shared__ int slen[];
    const int64_t row = (int64_t)blockIdx.x * sigma + threadIdx.x;
    const int len = (row < D) ? rowptr[row + 1] - rowptr[row] : -1;
    slen[threadIdx.x] = len;
    __syncthreads();
    int rank = 0;
    for (int j = 0; j < sigma; ++j) {
        const int lj = slen[j];
        rank += (lj > len) || (lj == len && j < (int)threadIdx.x);
    }
    slot_row[(int64_t)blockIdx.x * sigma + rank] = (row < D) ? (int)row : -1;
}

__global__ void k_sell_sizes(int64_t nslices, const int* __restrict__ rowptr, const int* __restrict__ slot_row,
                             int* __restrict__ sizes)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslices) return;
    int w = 0;
    for (int q = 0; q < 32; ++q) {
        const int r = slot_row[(s << 5) + q];
        if (r >= 0) w = max(w, rowptr[r + 1] - rowptr[r]);
    }
    sizes[s] = w << 5;
}

__global__ void k_sell_fill(int64_t D, int64_t nslices, const int* __restrict__ rowptr, const int* __restrict__ col,
                            const double* __restrict__ valJ, const int* __restrict__ slot_row, const int* __restrict__ sptr,
                            int* __restrict__ scol, double* __restrict__ sval, int* __restrict__ sdiag)
{
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= nslices) return;
    const int lane = threadIdx.x & 31;
    const int r = slot_row[(s << 5) + lane];
    const int base = sptr[s], w = (sptr[s + 1] - base) >> 5;
    int a = 0, len = 0;
    if (r >= 0) {
        a = rowptr[r];
        len = rowptr[r + 1] - a;
    }
    const int self = (r >= 0) ? r : (int)min((s << 5) + lane, D - 1);
    for (int j = 0; j < w; ++j) {
        int c = self;
        double v = 0.0;
        if (j < len) {
            c = col[a + j];
            v = valJ[a + j];
            if (c == r) sdiag[r] = base + (j << 5) + lane;
        }
        scol[base + (j << 5) + lane] = c;
        sval[base + (j << 5) + lane] = v;
    }
}

__global__ void k_sell_scale(int64_t n, const double* __restrict__ valJ, double cJ, double* __restrict__ valH)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        valH[e] = __dmul_rn(valJ[e], cJ);
}

__global__ void k_sell_diag(int64_t D, int n, const int* __restrict__ sdiag, const double* __restrict__ valJ,
                            const double* __restrict__ dU, double cJ, double cU, double cmu, double* __restrict__ valH)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D) return;
    const int p = sdiag[r];
    const double a = __dadd_rn(__dmul_rn(valJ[p], cJ), __dmul_rn(dU[r], cU));
    valH[p] = __dadd_rn(a, __dmul_rn(-(double)n, cmu));
}

static int build_sell(bh_ctx* ctx)
{
    if (!ctx->user_matrix) BH_TRY(bh_build_hamiltonian(ctx));
    const int64_t D = ctx->D;
    const int sigma = ctx->sell_sigma;  // 32 = no sorting beyond the slice itself
    const int64_t nwin = (D + sigma - 1) / sigma;
    const int64_t ns = nwin * (sigma / 32);
    int* d_sizes = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_sizes, sizeof(int) * ns));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_sell_ptr, sizeof(int) * (ns + 1)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_sell_row, sizeof(int) * ns * 32));
    k_sell_perm<<<(int)nwin, sigma, sizeof(int) * sigma, ctx->stream>>>(D, sigma, ctx->d_rowptr, ctx->d_sell_row);
    k_sell_sizes<<<nblocks(ns, 256), 256, 0, ctx->stream>>>(ns, ctx->d_rowptr, ctx->d_sell_row, d_sizes);
    ctx->launches += 2;
    int64_t total = 0;
    BH_TRY(bh_exclusive_scan(ctx, ns, d_sizes, ctx->d_sell_ptr, &total));
    cudaFree(d_sizes);
    if (total >= ((int64_t)1 << 31)) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "SELL copy has >= 2^31 entries");
    ctx->sell_nslices = ns;
    ctx->sell_entries = total;
    BH_CUDA(ctx, cudaMalloc(&ctx->d_sell_col, sizeof(int) * std::max<int64_t>(total, 1)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_sell_valJ, sizeof(double) * std::max<int64_t>(total, 1)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_sell_valH, sizeof(double) * std::max<int64_t>(total, 1)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_sell_diag, sizeof(int) * D));
    BH_CUDA(ctx, cudaMemsetAsync(ctx->d_sell_diag, 0, sizeof(int) * D, ctx->stream));
    k_sell_fill<<<nblocks(ns, 8), 256, 0, ctx->stream>>>(D, ns, ctx->d_rowptr, ctx->d_col, ctx->d_valJ, ctx->d_sell_row,
                                                         ctx->d_sell_ptr, ctx->d_sell_col, ctx->d_sell_valJ, ctx->d_sell_diag);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    ctx->sell_valid = ctx->sell_partial_valid = false;
    ctx->sell_cJ = ctx->sell_cU = ctx->sell_cmu = 0.0;
    return BH_OK;
}

int bh_materialise_sell(bh_ctx* ctx, double cJ, double cU, double cmu, int64_t max_entries, int64_t max_rows)
{
    const int64_t D = ctx->D;
    if (!ctx->d_sell_ptr) BH_TRY(build_sell(ctx));
    if ((ctx->sell_valid || (max_entries >= 0 && ctx->sell_partial_valid)) && ctx->sell_cJ == cJ && ctx->sell_cU == cU && ctx->sell_cmu == cmu) return BH_OK;
    // (the hybrid H.v only reads the slices of its stored rows: materialise just those)
    const int64_t ne = (max_entries >= 0) ? max_entries : ctx->sell_entries;
    const int64_t nr = (max_rows >= 0) ? max_rows : D;
    k_sell_scale<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ne, ctx->d_sell_valJ, cJ, ctx->d_sell_valH);
    k_sell_diag<<<nblocks(nr, 256), 256, 0, ctx->stream>>>(nr, ctx->n, ctx->d_sell_diag, ctx->d_sell_valJ, ctx->d_dU, cJ, cU,
                                                          cmu, ctx->d_sell_valH);
    ctx->launches += 2;
    BH_CUDA(ctx, cudaGetLastError());
    ctx->sell_cJ = cJ;
    ctx->sell_cU = cU;
    ctx->sell_cmu = cmu;
    ctx->sell_valid = (max_entries < 0);  // a partial materialisation does not serve the full stored kernel
    ctx->sell_partial_valid = true;
    return BH_OK;
}

// ---------------------------------------------------------------------------------------------
// exports (any ordering): row `pos` of the output = LEX row perm[pos], columns relabelled and re-sorted
// ---------------------------------------------------------------------------------------------
enum { EXPORT_JTERM = 0, EXPORT_HSUM = 1 };

__global__ void k_export_len(int64_t D, const int* __restrict__ rowptr, const int* __restrict__ perm, int mode,
                             int* __restrict__ len)
{
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= D) return;
    const int64_t k = perm ? perm[pos] : pos;
    len[pos] = rowptr[k + 1] - rowptr[k] - (mode == EXPORT_JTERM ? 1 : 0);
}

__global__ void k_export_rows(int64_t D, int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                              const double* __restrict__ valJ, const double* __restrict__ dU,
                              const int* __restrict__ perm, const int* __restrict__ inv, int mode, double cJ, double cU,
                              double cmu, const int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ val)
{
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= D) return;
    const int64_t k = perm ? perm[pos] : pos;
    int c[BH_MAX_ROW];
    double v[BH_MAX_ROW];
    int len = 0;
    for (int e = rowptr[k]; e < rowptr[k + 1]; ++e) {
        const int lc = col[e];
        double x;
        if (lc == (int)k) {
            if (mode == EXPORT_JTERM) continue;
            x = __dadd_rn(__dadd_rn(__dmul_rn(valJ[e], cJ), __dmul_rn(dU[k], cU)), __dmul_rn(-(double)n, cmu));
        } else {
            x = __dmul_rn(valJ[e], cJ);
        }
        const int nc = inv ? inv[lc] : lc;
        int p = len++;
        while (p > 0 && c[p - 1] > nc) {
            c[p] = c[p - 1];
            v[p] = v[p - 1];
            --p;
        }
        c[p] = nc;
        v[p] = x;
    }
    const int base = outer[pos];
    for (int i = 0; i < len; ++i) {
        inner[base + i] = c[i];
        val[base + i] = v[i];
    }
}

__global__ void k_export_diag(int64_t D, int n, const double* __restrict__ dU, const int* __restrict__ perm, int term,
                              double coef, int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ val)
{
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos > D) return;
    outer[pos] = (int)pos;
    if (pos == D) return;
    const int64_t k = perm ? perm[pos] : pos;
    inner[pos] = (int)pos;
    // src/hamiltonian.cpp:209 "U * value" and :229 "-mu * value"
    val[pos] = (term == BH_TERM_U) ? __dmul_rn(coef, dU[k]) : __dmul_rn(-coef, (double)n);
}

static int order_maps(bh_ctx* ctx, int order, const int** perm, const int** inv)
{
    *perm = *inv = nullptr;
    if (order == BH_ORDER_LEX) return BH_OK;
    BH_TRY(bh_ensure_orderings(ctx));
    if (order == BH_ORDER_TAG_SORTED) {
        *perm = ctx->d_perm_tag;
        *inv = ctx->d_inv_tag;
    } else {
        *perm = ctx->d_inv_tag;
        *inv = ctx->d_perm_tag;
    }
    return BH_OK;
}

static int export_matrix(bh_ctx* ctx, int mode, double cJ, double cU, double cmu, int order, int32_t* outer,
                         int32_t* inner, double* val)
{
    BH_TRY(bh_build_hamiltonian(ctx));
    const int64_t D = ctx->D;
    const int64_t nnz = (mode == EXPORT_JTERM) ? ctx->nnzJ : ctx->nnzH;
    const int *perm, *inv;
    BH_TRY(order_maps(ctx, order, &perm, &inv));
    int *d_len = nullptr, *d_outer = nullptr, *d_inner = nullptr;
    double* d_val = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_len, sizeof(int) * D));
    BH_CUDA(ctx, cudaMalloc(&d_outer, sizeof(int) * (D + 1)));
    BH_CUDA(ctx, cudaMalloc(&d_inner, sizeof(int) * std::max<int64_t>(nnz, 1)));
    BH_CUDA(ctx, cudaMalloc(&d_val, sizeof(double) * std::max<int64_t>(nnz, 1)));
    k_export_len<<<nblocks(D, 256), 256, 0, ctx->stream>>>(D, ctx->d_rowptr, perm, mode, d_len);
    BH_LAUNCHED(ctx);
    int64_t total = 0;
    BH_TRY(bh_exclusive_scan(ctx, D, d_len, d_outer, &total));
    if (total != nnz) return bh_fail(ctx, BH_ERR_STATE, "export: inconsistent entry count");
    k_export_rows<<<nblocks(D, 128), 128, 0, ctx->stream>>>(D, ctx->n, ctx->d_rowptr, ctx->d_col, ctx->d_valJ, ctx->d_dU,
                                                             perm, inv, mode, cJ, cU, cmu, d_outer, d_inner, d_val);
    BH_LAUNCHED(ctx);
    BH_D2H(ctx, outer, d_outer, sizeof(int) * (D + 1));
    if (nnz) {
        BH_D2H(ctx, inner, d_inner, sizeof(int) * nnz);
        BH_D2H(ctx, val, d_val, sizeof(double) * nnz);
    }
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_len);
    cudaFree(d_outer);
    cudaFree(d_inner);
    cudaFree(d_val);
    return BH_OK;
}

extern "C" int bh_term_nnz(bh_ctx* ctx, int term, int64_t* nnz)
{
    if (!ctx || !ctx->D || ctx->user_matrix || ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_term_nnz: call bh_setup first");
    if (!nnz || term < 0 || term > 2) return bh_fail(ctx, BH_ERR_ARG, "bh_term_nnz: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    BH_TRY(bh_build_hamiltonian(ctx));
    *nnz = (term == BH_TERM_J) ? ctx->nnzJ : ctx->D;
    return BH_OK;
}

extern "C" int bh_hamiltonian_nnz(bh_ctx* ctx, int64_t* nnz)
{
    if (!ctx || !ctx->D || ctx->user_matrix || ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_hamiltonian_nnz: call bh_setup first");
    if (!nnz) return bh_fail(ctx, BH_ERR_ARG, "bh_hamiltonian_nnz: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    BH_TRY(bh_build_hamiltonian(ctx));
    *nnz = ctx->nnzH;
    return BH_OK;
}

extern "C" int bh_term_csc(bh_ctx* ctx, int term, double coef, int order, int32_t* outer, int32_t* inner, double* val)
{
    if (!ctx || !ctx->D || ctx->user_matrix || ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_term_csc: call bh_setup first");
    if (term < 0 || term > 2 || order < 0 || order > 2 || !outer || !inner || !val)
        return bh_fail(ctx, BH_ERR_ARG, "bh_term_csc: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    if (term == BH_TERM_J) return export_matrix(ctx, EXPORT_JTERM, coef, 0, 0, order, outer, inner, val);
    const int64_t D = ctx->D;
    const int *perm, *inv;
    BH_TRY(order_maps(ctx, order, &perm, &inv));
    int *d_outer = nullptr, *d_inner = nullptr;
    double* d_val = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_outer, sizeof(int) * (D + 1)));
    BH_CUDA(ctx, cudaMalloc(&d_inner, sizeof(int) * D));
    BH_CUDA(ctx, cudaMalloc(&d_val, sizeof(double) * D));
    k_export_diag<<<nblocks(D + 1, 256), 256, 0, ctx->stream>>>(D, ctx->n, ctx->d_dU, perm, term, coef, d_outer, d_inner,
                                                                 d_val);
    BH_LAUNCHED(ctx);
    BH_D2H(ctx, outer, d_outer, sizeof(int) * (D + 1));
    BH_D2H(ctx, inner, d_inner, sizeof(int) * D);
    BH_D2H(ctx, val, d_val, sizeof(double) * D);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_outer);
    cudaFree(d_inner);
    cudaFree(d_val);
    return BH_OK;
}

extern "C" int bh_hamiltonian_csc(bh_ctx* ctx, double cJ, double cU, double cmu, int order, int32_t* outer,
                                  int32_t* inner, double* val)
{
    if (!ctx || !ctx->D || ctx->user_matrix || ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_hamiltonian_csc: call bh_setup first");
    if (order < 0 || order > 2 || !outer || !inner || !val) return bh_fail(ctx, BH_ERR_ARG, "bh_hamiltonian_csc: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    return export_matrix(ctx, EXPORT_HSUM, cJ, cU, cmu, order, outer, inner, val);
}

// ---------------------------------------------------------------------------------------------
// user matrix (the literal Op::IRLM_eigen(SparseMatrix O, ..) seam, reference src/operator.cpp:22-33)
// ---------------------------------------------------------------------------------------------
extern "C" int bh_load_matrix(bh_ctx* ctx, int64_t D, const int32_t* outer, const int32_t* inner, const double* val)
{
    if (!ctx) return BH_ERR_ARG;
    if (D < 1 || D >= ((int64_t)1 << 31) || !outer || !inner || !val) return bh_fail(ctx, BH_ERR_ARG, "bh_load_matrix: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nnz = outer[D];
    if (outer[0] != 0 || nnz < 0) return bh_fail(ctx, BH_ERR_ARG, "bh_load_matrix: outer index array must start at 0");
    bh_release_system(ctx);
    ctx->user_matrix = true;
    ctx->D = D;
    ctx->row0 = 0;
    ctx->nloc = D;
    ctx->ld = (D + 31) / 32 * 32;
    ctx->nnzH = nnz;
    ctx->nnzJ = nnz;
    BH_CUDA(ctx, cudaMalloc(&ctx->d_rowptr, sizeof(int) * (D + 1 + 64)));
    BH_CUDA(ctx, cudaMemsetAsync(ctx->d_rowptr, 0, sizeof(int) * (D + 1 + 64), ctx->stream));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_col, sizeof(int) * (nnz + 8)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_valJ, sizeof(double) * (nnz + 8)));
    BH_H2D(ctx, ctx->d_rowptr, outer, sizeof(int) * (D + 1));
    if (nnz) {
        BH_H2D(ctx, ctx->d_col, inner, sizeof(int) * nnz);
        BH_H2D(ctx, ctx->d_valJ, val, sizeof(double) * nnz);
    }
    BH_TRY(build_sell(ctx));
    BH_CUDA(ctx, cudaMemcpyAsync(ctx->d_sell_valH, ctx->d_sell_valJ, sizeof(double) * ctx->sell_entries,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->sell_valid = true;
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // the CSR staging copy is no longer needed
    cudaFree(ctx->d_col); ctx->d_col = nullptr;
    cudaFree(ctx->d_valJ); ctx->d_valJ = nullptr;
    return BH_OK;
}
