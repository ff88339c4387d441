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
