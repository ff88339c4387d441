// device_utils.cuh -- device helpers shared by the kernels (packed Fock states, ranking, reductions).
#pragma once

#include "bh_internal.h"

// Occupation of site i in a packed state (4 bits per site, site 0 in the low nibble).
__device__ __forceinline__ int bh_occ(uint64_t s, int i) { return (int)((s >> (4 * i)) & 0xFull); }

// Cooperative copy of the tables into shared memory (one BhTables per CTA).
__device__ __forceinline__ void bh_stage_tables(BhTables* dst, const BhTables* __restrict__ src)
{
    const int words = (int)(sizeof(BhTables) / sizeof(int));
    const int* s = reinterpret_cast<const int*>(src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = s[i];
    __syncthreads();
}

// Prefix arrays for O(1) ranking of any single-boson move (see DESIGN.md "incremental rank"):
//   with R_q = bosons on sites > q,
//   dn[s] = sum_{q<s} f_q(R_q - 1) - f_q(R_q)   -> move src -> dst with dst < src: rank += dn[src] - dn[dst]
//   up[s] = sum_{q<s} f_q(R_q + 1) - f_q(R_q)   -> move src -> dst with dst > src: rank += up[dst] - up[src]
template <int M>
__device__ __forceinline__ void bh_rank_prefix(const BhTables& t, uint64_t s, int (&dn)[M], int (&up)[M])
{
    int R = t.n;
    int adn = 0, aup = 0;
#pragma unroll
    for (int q = 0; q < M; ++q) {
        dn[q] = adn;
        up[q] = aup;
        R -= bh_occ(s, q);
        if (q < M - 1) {
            const int f0 = t.f[q][R];
            adn += (R >= 1 ? t.f[q][R - 1] : f0) - f0;
            aup += t.f[q][R + 1] - f0;
        }
    }
}

// Descending-lexicographic rank of a packed state (m sites).
__device__ __forceinline__ int bh_rank_of(const BhTables& t, uint64_t s)
{
    int R = t.n, r = 0;
    for (int q = 0; q < t.m - 1; ++q) {
        R -= bh_occ(s, q);
        r += t.f[q][R];
    }
    return r;
}

// Off-diagonal part of one row of a nearest-neighbour chain, sum_hops sqrt((n_dst + 1) n_src) x[source], in a single sweep
// over the sites (see k_hv_free_chain in hv.cu; the caller multiplies by -2J).  x is read with ordinary loads: the callers
// in small.cu apply H repeatedly inside ONE kernel to vectors that the same CTA wrote a moment ago.
template <int M, bool CLOSED>
__device__ __forceinline__ double bh_chain_row_sum(const BhTables& t, uint64_t s, int kk, const double* x)
{
    const int n0 = bh_occ(s, 0);
    int R = t.n - n0, nprev = n0, tdn = 0, tup = 0;
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < M - 1; ++q) {
        const int nnext = bh_occ(s, q + 1);
        const int2 gh = t.gh[q][R];  // .x: boson moves q+1 -> q, .y: q -> q+1
        const double xa = nnext ? x[kk + gh.x] : 0.0;
        const double xb = nprev ? x[kk + gh.y] : 0.0;
        acc = fma(t.sq[(nprev + 1) * nnext], xa, acc);
        acc = fma(t.sq[(nnext + 1) * nprev], xb, acc);
        tdn += gh.x;
        tup += gh.y;
        R -= nnext;
        nprev = nnext;
    }
    if (CLOSED) {
        const int nl = nprev;  // occupation of the last site
        const double xa = nl ? x[kk + tdn] : 0.0;  // M-1 -> 0
        const double xb = n0 ? x[kk + tup] : 0.0;  // 0 -> M-1
        acc = fma(t.sq[(n0 + 1) * nl], xa, acc);
        acc = fma(t.sq[(nl + 1) * n0], xb, acc);
    }
    return acc;
}

__device__ __forceinline__ double bh_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum (all threads get the result in thread 0 only); scratch has >= 32 doubles.
__device__ __forceinline__ double bh_block_sum(double v, double* scratch)
{
    v = bh_warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.0;
    if (wid == 0) v = bh_warp_sum(v);
    return v;
}
