// hv_split.cu -- K4, "split" matrix-free H.v for chains (the reference CLI's only geometry, src/analysis.cpp:219-220).
//
// Replaces SparseGenMatProd::perform_op (external/spectra/include/Spectra/MatOp/SparseGenMatProd.h:81-86) on the
// Hamiltonian of src/hamiltonian.cpp:170-232 without storing it and without decoding a Fock state per row.
// The chain is cut after site p-1 (hv_split_tables.h): with (P, S) = (prefix, suffix occupations),
// rank(P,S) = off(P) + sufrank(S), the vector of the sector "R bosons in the suffix" is a matrix X_R[P][S] with
// contiguous rows, and H = (prefix bonds) (x) 1 + 1 (x) (suffix bonds) + cut bond + periodic bond + diagonal.
//
// Mapping: a warp owns 32 consecutive suffixes x G prefixes of one sector (G accumulators per lane).
//   * suffix hops:  one packed (sufrank, amplitude code) word per hop, loaded coalesced ONCE per lane and reused
//                   for the G prefixes; the gathers x[off(P) + sufrank'] stay inside the row (L1/L2 local);
//   * prefix hops:  (off(P'), amplitude) is uniform over the warp (shared-memory broadcast), the x reads are coalesced;
//   * cut / periodic bond: product of a prefix factor (uniform) and a suffix factor (per lane), 4 terms per row;
//   * diagonal:     U (dU(P) + dU(S)) - mu n, bit-identical to the stored value (integers below 2^53).
// ~5 instructions per hop instead of ~22 for the state-decoding chain kernel (hv.cu, k_hv_free_chain).
// A CTA = 8 warps = (nx suffix chunks) x (ny prefix groups) of one sector; CTAs are issued sector by sector, so at
// m = n = 14 the working set of a moment (three neighbouring sectors, <= 70 MB) stays in the 126 MB L2 and every
// element of x is fetched from HBM once: traffic = 16 B per row, the algorithmic figure of SURVEY.md section 8(d).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bh_internal.h"
#include "hv_split_tables.h"

struct SplitDev {
    int n, closed, WP, WS, rec_bytes;
    uint32_t nitems;
    int ablate;  // experiments only (env BH_SPLIT_ABLATE): 1 skip suffix bonds, 2 skip prefix bonds, 4 skip cut/periodic bonds
    const unsigned char* prec;
    const uint32_t* sinfo;
    const uint4* scross;
    const uint32_t* snbr;
    SplitSector sec[BH_SPLIT_MAX_SECTORS + 1];
};

struct bh_split_state {
    SplitDev dev;
    int G = 0, UJ = 0, p = 0;
    size_t smem = 0;
    size_t table_bytes = 0;
};

#ifndef SPLIT_MINB
#define SPLIT_MINB 2
#endif

// x element at BYTE offset b (32-bit: D < 2^29 for every supported system, m <= 16 and n <= 15 give D <= C(30,15))
__device__ __forceinline__ double ldx(const double* __restrict__ x, uint32_t b)
{
    return __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(x) + b));
}

// One warp: 32 suffixes (lane) x up to G prefixes.  FULL = all G prefixes present (no per-prefix guards in the loops).
// UJ hops are processed together so that UJ * G independent gathers are in flight per lane (the kernel is latency-bound
// otherwise: ncu long-scoreboard stalls on the first use of every loaded value).
template <int G, int UJ, bool FULL>
__device__ __forceinline__ void split_warp(const SplitDev& T, const double* sq, const unsigned char* recb, int gcount, int R,
                                           uint32_t nS, uint32_t nSpad, uint32_t sbase, uint32_t S, double cJ, double cU,
                                           double cmu, const double* __restrict__ x, double* __restrict__ y,
                                           const BhEpilogue& ep)
{
    const bool valid = S < nS;
    const uint32_t Sb = (valid ? S : nS - 1) * 8u;
    const uint32_t* nb = T.snbr + (size_t)sbase * T.WS + S;
    const int WS = T.WS;
    // first block of suffix-hop words, issued before anything depends on them
    uint32_t vcur[UJ];
#pragma unroll
    for (int u = 0; u < UJ; ++u) vcur[u] = (u < WS) ? __ldg(nb + (size_t)u * nSpad) : 0u;
    const uint32_t si = __ldg(T.sinfo + sbase + S);
    const uint4 cr = __ldg(T.scross + sbase + S);
    const uint32_t rb = T.rec_bytes;

    uint32_t ob[G];  // byte offset of row P_g
    double acc[G], xv[G];
    int pmax = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        acc[g] = 0.0;
        xv[g] = 0.0;
        ob[g] = 0;
        if (FULL || g < gcount) {
            const SplitPrefixHdr* h = reinterpret_cast<const SplitPrefixHdr*>(recb + g * rb);
            ob[g] = h->off;
            xv[g] = ldx(x, ob[g] + Sb);
            pmax = max(pmax, (int)((h->info >> 8) & 255));
        }
    }

    // ---- prefix bonds: A (x) 1 (padding entries point at the own row with amplitude 0) ----
    {
        const unsigned char* nbp = recb + sizeof(SplitPrefixHdr);
        if (!(T.ablate & 2))
        for (int j0 = 0; j0 < pmax; j0 += UJ) {
            double val[UJ][G], a[UJ][G];
#pragma unroll
            for (int u = 0; u < UJ; ++u)
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    val[u][g] = 0.0;
                    a[u][g] = 0.0;
                    if ((FULL || g < gcount) && j0 + u < pmax) {
                        const uint4 e = *reinterpret_cast<const uint4*>(nbp + g * rb + (j0 + u) * sizeof(SplitNbr));
                        a[u][g] = __hiloint2double((int)e.w, (int)e.z);
                        val[u][g] = ldx(x, e.x + Sb);  // e.x = byte offset of row P'
                    }
                }
#pragma unroll
            for (int u = 0; u < UJ; ++u)
#pragma unroll
                for (int g = 0; g < G; ++g) acc[g] = fma(a[u][g], val[u][g], acc[g]);
        }
    }

    const int np = si & 15, nl = (si >> 4) & 15, scnt = (si >> 8) & 255, dUs = si >> 16;
    // ---- suffix bonds: 1 (x) B ----
    {
        const int smax = (T.ablate & 1) ? 0 : __reduce_max_sync(0xffffffffu, scnt);
        for (int j0 = 0; j0 < smax; j0 += UJ) {
            uint32_t vnext[UJ];
#pragma unroll
            for (int u = 0; u < UJ; ++u) vnext[u] = (j0 + UJ + u < smax) ? __ldg(nb + (size_t)(j0 + UJ + u) * nSpad) : 0u;
            double val[UJ][G], a[UJ];
#pragma unroll
            for (int u = 0; u < UJ; ++u) {
                const bool on = j0 + u < smax;
                a[u] = on ? sq[vcur[u] >> 24] : 0.0;
                const uint32_t ib = (vcur[u] & 0xffffffu) * 8u;
#pragma unroll
                for (int g = 0; g < G; ++g) val[u][g] = ((FULL || g < gcount) && on) ? ldx(x, ob[g] + ib) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UJ; ++u) {
#pragma unroll
                for (int g = 0; g < G; ++g) acc[g] = fma(a[u], val[u][g], acc[g]);
                vcur[u] = vnext[u];
            }
        }
    }

    // ---- cut bond (p-1, p) and periodic bond (m-1, 0); then the diagonal and the fused epilogue ----
    const double shift = __dmul_rn(-(double)T.n, cmu);
    const double twoJ = 2.0 * cJ;
    const uint32_t cx = cr.x * 8u, cy = cr.y * 8u, cz = cr.z * 8u, cw = cr.w * 8u;
    double c0[G], c1[G], c2[G], c3[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        c0[g] = c1[g] = c2[g] = c3[g] = 0.0;
        if ((FULL || g < gcount) && !(T.ablate & 4)) {
            const SplitPrefixHdr* h = reinterpret_cast<const SplitPrefixHdr*>(recb + g * rb);
            const uint4 o = *reinterpret_cast<const uint4*>(h);  // off_cu, off_cd, off_wu, off_wd (byte offsets)
            const uint32_t info = h->info;
            const int n0 = info & 15, nq = (info >> 4) & 15;
            if (R >= 1) {
                if (np) c0[g] = sq[(nq + 1) * np] * ldx(x, o.x + cx);               // boson moves p -> p-1
                if (T.closed && nl) c2[g] = sq[(n0 + 1) * nl] * ldx(x, o.z + cz);  // boson moves m-1 -> 0
            }
            if (nq) c1[g] = sq[(np + 1) * nq] * ldx(x, o.y + cy);                  // boson moves p-1 -> p
            if (T.closed && n0) c3[g] = sq[(nl + 1) * n0] * ldx(x, o.w + cw);      // boson moves 0 -> m-1
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
        if ((FULL || g < gcount) && valid) {
            const SplitPrefixHdr* h = reinterpret_cast<const SplitPrefixHdr*>(recb + g * rb);
            const int dUp = h->info >> 16;
            const double a = acc[g] + ((c0[g] + c1[g]) + (c2[g] + c3[g]));
            const size_t l = (size_t)(ob[g] >> 3) + S;
            const double diag = __dadd_rn(__dmul_rn((double)(dUp + dUs), cU), shift);
            double out = ep.s1 * (diag * xv[g] - twoJ * a);
            if (ep.s2 != 0.0) out = fma(ep.s2, xv[g], out);
            if (ep.z) out = fma(ep.s3, __ldcs(ep.z + l), out);
            __stcs(y + l, out);
        }
}

template <int G, int UJ, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_hv_split(const __grid_constant__ SplitDev T, const BhTables* __restrict__ gtab, double cJ, double cU, double cmu,
           const double* __restrict__ x, double* __restrict__ y, BhEpilogue ep)
{
    extern __shared__ uint4 s_rec[];  // the CTA's prefix records (ny * G of them), row offsets converted to bytes
    __shared__ double sq[256];
    const int tid = threadIdx.x;
    const uint32_t item = blockIdx.x;
    int R = 0;
    while (R < T.n && item >= T.sec[R + 1].item_first) ++R;
    const uint32_t nS = T.sec[R].nS, nSpad = T.sec[R].nSpad, nP = T.sec[R].nP, sbase = T.sec[R].sbase;
    const uint32_t nx = T.sec[R].nx, ny = T.sec[R].ny, ncb = T.sec[R].ncb;
    const uint32_t local = item - T.sec[R].item_first;
    const uint32_t gb = local / ncb, cb = local - gb * ncb;
    const uint32_t pg0 = gb * ny * G;  // first prefix of this CTA inside the sector
    const uint32_t npre = min(ny * (uint32_t)G, nP - pg0);
    {
        const uint4* src = reinterpret_cast<const uint4*>(T.prec + (size_t)(T.sec[R].pfirst + pg0) * T.rec_bytes);
        const int nvec = (int)npre * (T.rec_bytes >> 4);
        for (int i = tid; i < nvec; i += 256) s_rec[i] = __ldg(src + i);
        sq[tid] = gtab->sq[tid];
    }
    __syncthreads();
    const uint32_t warp = tid >> 5, lane = tid & 31;
    const uint32_t wy = warp / nx, wx = warp - wy * nx;
    const uint32_t chunk = cb * nx + wx;
    if (chunk * 32 >= nSpad) return;
    const int gcount = min(G, (int)npre - (int)(wy * G));
    if (gcount <= 0) return;
    const uint32_t S = chunk * 32 + lane;
    const unsigned char* recb = reinterpret_cast<const unsigned char*>(s_rec) + (size_t)(wy * G) * T.rec_bytes;
    if (gcount == G)
        split_warp<G, UJ, true>(T, sq, recb, gcount, R, nS, nSpad, sbase, S, cJ, cU, cmu, x, y, ep);
    else
        split_warp<G, UJ, false>(T, sq, recb, gcount, R, nS, nSpad, sbase, S, cJ, cU, cmu, x, y, ep);
}

typedef void (*hv_split_fn)(const SplitDev, const BhTables*, double, double, double, const double*, double*, BhEpilogue);

static hv_split_fn split_kernel(int G, int UJ)
{
    switch (G * 16 + UJ) {
        case 2 * 16 + 2: return k_hv_split<2, 2, 5>;
        case 2 * 16 + 4: return k_hv_split<2, 4, 4>;
        case 2 * 16 + 6: return k_hv_split<2, 6, 4>;
        case 4 * 16 + 1: return k_hv_split<4, 1, 3>;
        case 4 * 16 + 2: return k_hv_split<4, 2, 3>;
        case 4 * 16 + 3: return k_hv_split<4, 3, 3>;
        case 4 * 16 + 4: return k_hv_split<4, 4, 2>;
        case 8 * 16 + 1: return k_hv_split<8, 1, 2>;
        case 8 * 16 + 2: return k_hv_split<8, 2, 2>;
        case 16 * 16 + 1: return k_hv_split<16, 1, 1>;
    }
    return nullptr;
}

void bh_split_release(bh_ctx* ctx)
{
    bh_split_state* st = static_cast<bh_split_state*>(ctx->split);
    if (!st) return;
    cudaFree(const_cast<unsigned char*>(st->dev.prec));
    cudaFree(const_cast<uint32_t*>(st->dev.sinfo));
    cudaFree(const_cast<uint4*>(st->dev.scross));
    cudaFree(const_cast<uint32_t*>(st->dev.snbr));
    delete st;
    ctx->split = nullptr;
}

bool bh_split_supported(const bh_ctx* ctx)
{
    return !ctx->user_matrix && !ctx->partitioned && ctx->h_tab.chain != 0 && ctx->m >= 3 && ctx->n >= 1 &&
           ctx->n + 1 <= BH_SPLIT_MAX_SECTORS && ctx->D < ((int64_t)1 << 29);
}

static int ensure_tables(bh_ctx* ctx)
{
    if (ctx->split) return BH_OK;
    int G = ctx->split_G;
    const int UJ = ctx->split_UJ;
    if (!split_kernel(G, UJ)) return bh_fail(ctx, BH_ERR_ARG, "unsupported BH_SPLIT_G / BH_SPLIT_UJ combination");
    int p = ctx->split_p > 0 ? ctx->split_p : ctx->m / 2;
    p = std::min(std::max(p, 1), ctx->m - 1);
    SplitTables T;
    try {
        bh_split_build(ctx->m, ctx->n, p, G, ctx->h_tab.chain == 2, &ctx->h_tab.f[0][0], BH_MAX_BOSONS + 3, T, ctx->split_nx);
    } catch (const std::exception& e) {
        return bh_fail(ctx, BH_ERR_STATE, std::string("split H.v tables: ") + e.what());
    }
    // the kernel addresses x with 32-bit BYTE offsets: convert every row offset of the prefix records
    for (uint32_t i = 0; i < T.NP; ++i) {
        unsigned char* rec = T.prec.data() + (size_t)i * T.rec_bytes;
        SplitPrefixHdr* h = reinterpret_cast<SplitPrefixHdr*>(rec);
        SplitNbr* nb = reinterpret_cast<SplitNbr*>(rec + sizeof(SplitPrefixHdr));
        h->off *= 8u; h->off_cu *= 8u; h->off_cd *= 8u; h->off_wu *= 8u; h->off_wd *= 8u;
        for (int j = 0; j < T.WP; ++j) nb[j].off *= 8u;
    }
    bh_split_state* st = new bh_split_state();
    st->G = G;
    st->UJ = UJ;
    st->p = p;
    SplitDev& d = st->dev;
    std::memset(&d, 0, sizeof(d));
    d.ablate = getenv("BH_SPLIT_ABLATE") ? atoi(getenv("BH_SPLIT_ABLATE")) : 0;
    d.n = T.n; d.closed = T.closed; d.WP = T.WP; d.WS = T.WS; d.rec_bytes = T.rec_bytes; d.nitems = T.nitems;
    std::memcpy(d.sec, T.sec, sizeof(d.sec));
    ctx->split = st;  // from here bh_split_release frees whatever was allocated
    void* ptr = nullptr;
    const size_t b_prec = std::max<size_t>(T.prec.size(), 16), b_info = std::max<size_t>(T.sinfo.size() * 4, 16),
                 b_cross = std::max<size_t>(T.scross.size() * 4, 16), b_nbr = std::max<size_t>(T.snbr.size() * 4, 16);
    BH_CUDA(ctx, cudaMalloc(&ptr, b_prec)); d.prec = static_cast<unsigned char*>(ptr);
    BH_CUDA(ctx, cudaMalloc(&ptr, b_info)); d.sinfo = static_cast<uint32_t*>(ptr);
    BH_CUDA(ctx, cudaMalloc(&ptr, b_cross)); d.scross = static_cast<uint4*>(ptr);
    BH_CUDA(ctx, cudaMalloc(&ptr, b_nbr)); d.snbr = static_cast<uint32_t*>(ptr);
    BH_H2D(ctx, const_cast<unsigned char*>(d.prec), T.prec.data(), T.prec.size());
    BH_H2D(ctx, const_cast<uint32_t*>(d.sinfo), T.sinfo.data(), T.sinfo.size() * 4);
    BH_H2D(ctx, const_cast<uint4*>(d.scross), T.scross.data(), T.scross.size() * 4);
    if (!T.snbr.empty()) BH_H2D(ctx, const_cast<uint32_t*>(d.snbr), T.snbr.data(), T.snbr.size() * 4);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors die with this scope
    st->table_bytes = b_prec + b_info + b_cross + b_nbr;
    st->smem = (size_t)8 * G * T.rec_bytes;
    if (st->smem > 200 * 1024) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "split H.v: prefix records do not fit in shared memory");
    if (st->smem > 48 * 1024)
        BH_CUDA(ctx, cudaFuncSetAttribute(split_kernel(G, UJ), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem));
    if (getenv("BH_SPLIT_VERBOSE"))
        fprintf(stderr, "[bh] split H.v: m=%d n=%d p=%d G=%d items=%u prefixes=%u suffixes(padded)=%u tables=%.1f MB smem=%zu\n", ctx->m,
                ctx->n, p, G, T.nitems, T.NP, T.NSpad, st->table_bytes / 1e6, st->smem);
    return BH_OK;
}

int bh_launch_hv_split(bh_ctx* ctx, double cJ, double cU, double cmu, const double* x, double* y, const BhEpilogue& ep)
{
    if (!bh_split_supported(ctx)) return bh_fail(ctx, BH_ERR_STATE, "split H.v needs a chain on an unpartitioned context");
    BH_TRY(ensure_tables(ctx));
    bh_split_state* st = static_cast<bh_split_state*>(ctx->split);
    split_kernel(st->G, st->UJ)<<<st->dev.nitems, 256, st->smem, ctx->stream>>>(st->dev, ctx->d_tab, cJ, cU, cmu, x, y, ep);
    BH_LAUNCHED(ctx);
    return BH_OK;
}
