// lanczos.cu -- K5/K6: FP64 thick-restart Lanczos on the GPU with Spectra's parameters and stopping rule.
//
// Replaces Op::IRLM_eigen -> Spectra::GenEigsSolver (reference src/operator.cpp:22-33) for the symmetric H.
// The CPU statement followed is Spectra's symmetric solver: start vector v = A v0 / |A v0| with v0 the
// LCG(seed 0) vector (HermEigsBase.h:331-336, LinAlg/Arnoldi.h:132-180), the Lanczos step of
// LinAlg/Lanczos.h:59-184 (three-term recurrence, then a full re-orthogonalisation correction of the
// residual against every basis vector), Ritz pairs of the projected matrix, convergence when
// |last-row Ritz component| * |f| < tol * max(eps^(2/3), |theta|) (HermEigsBase.h:152-169), restart size
// nev + min(nconv, (ncv - nev)/2) (HermEigsBase.h:172-196).  The implicit restart with exact shifts
// (HermEigsBase.h:102-148) is done in its mathematically equivalent thick-restart form: keep the first k
// Ritz vectors V <- V Y_k (K6), the projected matrix becomes diag(theta) plus one coupling row.
//
// Device design: all vectors and the recurrence scalars (alpha, beta, re-orthogonalisation coefficients)
// stay in HBM; a Lanczos step is 7 launches with no host synchronisation (each reduction kernel finishes
// in its last-arriving block, deterministically); the host reads the ncv-sized scalar arrays once per
// restart, solves the <=128 x 128 projected problem and uploads the Ritz coefficients.
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <limits>

#include "device_utils.cuh"

#define NC (BH_MAX_NCV + 2)
// layout of ctx->d_scal (doubles)
#define S_ALPHA 0            // alpha[i]            = T(i,i)
#define S_BETA (NC)          // beta[i+1] = |f| after step i ; beta[0] = |A v0|
#define S_OFFD (2 * NC)      // offd[i]             = T(i,i-1) including the re-orthogonalisation correction
#define S_VF (3 * NC)        // coefficients of the last re-orthogonalisation pass
#define S_FLAG (4 * NC)      // [0] breakdown flag
#define S_TOTAL (4 * NC + 8)

#define VEC_THREADS 256
#define GT_CH 8  // columns accumulated per sweep of the transposed product

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }
namespace cg = cooperative_groups;

__device__ __forceinline__ bool bh_last_block(unsigned int* counter)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
    __syncthreads();
    return is_last;
}

// LCG(16807) sequence starting at element `first` of the global stream (fresh vectors after a breakdown)
__global__ void k_lcg_seq(int64_t first, int64_t n, double* __restrict__ out)
{
    const unsigned long long M = 2147483647ull;
    const int64_t start = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 64;
    if (start >= n) return;
    unsigned long long seed = 1, base = 16807ull;
    for (unsigned long long p = (unsigned long long)(first + start); p; p >>= 1) {
        if (p & 1) seed = seed * base % M;
        base = base * base % M;
    }
    for (int64_t i = start; i < min(start + 64, n); ++i) {
        seed = seed * 16807ull % M;
        out[i] = (double)(long long)seed / 2147483647.0 - 0.5;
    }
}

// |v| -> scal[dst]  (used once for |A v0|)
__global__ void __launch_bounds__(VEC_THREADS)
k_norm(int64_t D, const double* __restrict__ v, double* __restrict__ scal, int dst, double* __restrict__ part,
       unsigned int* __restrict__ counter, int raw = 0)
{
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < D; r += (int64_t)gridDim.x * blockDim.x)
        acc += v[r] * v[r];
    acc = bh_block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
    if (bh_last_block(counter)) {
        double t = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += part[b];
        t = bh_block_sum(t, scratch);
        if (threadIdx.x == 0) scal[dst] = raw ? t : sqrt(t);  // raw: local sum of squares, all-reduced by the caller
    }
}

// V_i = f / beta  (beta = scal[S_BETA + i]); a vanishing beta raises the breakdown flag
__global__ void __launch_bounds__(VEC_THREADS)
k_scale(int64_t D, const double* __restrict__ f, double* __restrict__ vi, double* __restrict__ scal, int i, double thresh)
{
    const double beta = scal[S_BETA + i];
    double inv = 0.0;
    if (beta > thresh)
        inv = 1.0 / beta;
    else if (blockIdx.x == 0 && threadIdx.x == 0 && scal[S_FLAG] == 0.0) {
        scal[S_FLAG] = 1.0;          // breakdown: the Krylov space became invariant ...
        scal[S_FLAG + 1] = (double)i;  // ... before basis column i could be formed
    }
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < D; r += (int64_t)gridDim.x * blockDim.x)
        vi[r] = f[r] * inv;
}

// w -= offd * V_{i-1} (optional) ; alpha_i = V_i . w
__global__ void __launch_bounds__(VEC_THREADS)
k_local_alpha(int64_t D, double* __restrict__ w, const double* __restrict__ vprev, const double* __restrict__ vi,
              double* __restrict__ scal, int i, int subtract, double* __restrict__ part, unsigned int* __restrict__ counter)
{
    __shared__ double scratch[32];
    const double b = subtract ? scal[S_BETA + i] : 0.0;
    double acc = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < D; r += (int64_t)gridDim.x * blockDim.x) {
        double wr = w[r];
        if (subtract) {
            wr -= b * vprev[r];
            w[r] = wr;
        }
        acc += vi[r] * wr;
    }
    acc = bh_block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
    if (bh_last_block(counter)) {
        double t = 0.0;
        for (int bb = threadIdx.x; bb < (int)gridDim.x; bb += blockDim.x) t += part[bb];
        t = bh_block_sum(t, scratch);
        if (threadIdx.x == 0) {
            scal[S_ALPHA + i] = t;
            scal[S_OFFD + i] = b;
        }
    }
}

// f = w - alpha_i V_i ; beta_{i} = |f|
__global__ void __launch_bounds__(VEC_THREADS)
k_resid_norm(int64_t D, const double* __restrict__ w, const double* __restrict__ vi, double* __restrict__ f,
             double* __restrict__ scal, int i, double* __restrict__ part, unsigned int* __restrict__ counter)
{
    __shared__ double scratch[32];
    const double a = scal[S_ALPHA + i];
    double acc = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < D; r += (int64_t)gridDim.x * blockDim.x) {
        const double fr = w[r] - a * vi[r];
        f[r] = fr;
        acc += fr * fr;
    }
    acc = bh_block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
    if (bh_last_block(counter)) {
        double t = 0.0;
        for (int bb = threadIdx.x; bb < (int)gridDim.x; bb += blockDim.x) t += part[bb];
        t = bh_block_sum(t, scratch);
        if (threadIdx.x == 0) scal[S_BETA + i + 1] = sqrt(t);
    }
}


// ---- full re-orthogonalisation of the residual, in column blocks of at most GT_CH basis vectors ----------
// For each block B of columns (block modified Gram-Schmidt): k_gemv_t computes c_B = V_B^T f, k_gemv_n_norm
// applies f -= V_B c_B.  A block is small enough (GT_CH * 8 * D bytes) to stay in L2 between the two kernels, so
// the basis is read from HBM once per Lanczos step rather than twice.  Rows are handled in pairs (16-byte loads).
//
// c[col0 .. col0+cnt) = V[:, col0 .. col0+cnt)^T f ; then alpha_i += c[i], offd_i += c[i-1]  (Lanczos.h:150-171)
// (cnt may exceed GT_CH: the columns are then swept in chunks of GT_CH with f re-read from L1/L2 -- used for
//  small systems whose whole basis is L2-resident, where fewer launches matter more than HBM passes)
template <bool MULTI>
__global__ void __launch_bounds__(VEC_THREADS)
k_gemv_t(int64_t D, int64_t ld, const double* __restrict__ V, int col0, int cnt, const double* __restrict__ f,
         double* __restrict__ scal, int i, int fix_offd, double* __restrict__ part, unsigned int* __restrict__ counter,
         int raw = 0)
{
    __shared__ double red[VEC_THREADS / 32][GT_CH];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t npair = (D + 1) >> 1;  // ld is even and the padding is zero
    const double2* f2 = reinterpret_cast<const double2*>(f);
    const int64_t ld2 = ld >> 1;
    const int pstride = MULTI ? NC : GT_CH;
    const int ncc = MULTI ? cnt : 1;  // single-chunk instantiation: cnt <= GT_CH, loop body runs once
    for (int cc = 0; cc < ncc; cc += GT_CH) {
        const int nc = min(GT_CH, cnt - cc);
        const double2* V2 = reinterpret_cast<const double2*>(V + (int64_t)(col0 + cc) * ld);
        double acc[GT_CH];
#pragma unroll
        for (int j = 0; j < GT_CH; ++j) acc[j] = 0.0;
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < npair; r += (int64_t)gridDim.x * blockDim.x) {
            const double2 fr = f2[r];
            double2 v[GT_CH];
#pragma unroll
            for (int j = 0; j < GT_CH; ++j) v[j] = (j < nc) ? V2[(int64_t)j * ld2 + r] : make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < GT_CH; ++j) acc[j] = fma(v[j].x, fr.x, fma(v[j].y, fr.y, acc[j]));
        }
#pragma unroll
        for (int j = 0; j < GT_CH; ++j) acc[j] = bh_warp_sum(acc[j]);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < GT_CH; ++j) red[wid][j] = acc[j];
        }
        __syncthreads();
        if (threadIdx.x < GT_CH) {
            double t = 0.0;
            for (int w = 0; w < VEC_THREADS / 32; ++w) t += red[w][threadIdx.x];
            part[(int64_t)blockIdx.x * pstride + cc + threadIdx.x] = t;
        }
    }
    if (bh_last_block(counter)) {
        // one warp per column: lanes stride over the blocks' partial sums, fixed-order tree
        for (int j = wid; j < cnt; j += VEC_THREADS / 32) {
            double t = 0.0;
            for (int b = lane; b < (int)gridDim.x; b += 32) t += part[(int64_t)b * pstride + j];
            t = bh_warp_sum(t);
            if (lane == 0) {
                const int c = col0 + j;
                scal[S_VF + c] = t;
                if (!raw && c == i) scal[S_ALPHA + i] += t;
                if (!raw && fix_offd && c == i - 1) scal[S_OFFD + i] += t;
            }
        }
    }
}

// f -= V[:, col0 .. col0+cnt) c ; beta_i = |f|
template <bool MULTI>
__global__ void __launch_bounds__(VEC_THREADS)
k_gemv_n_norm(int64_t D, int64_t ld, const double* __restrict__ V, int col0, int cnt, double* __restrict__ f,
              double* __restrict__ scal, int i, double* __restrict__ part, unsigned int* __restrict__ counter, int raw = 0)
{
    __shared__ double scratch[32];
    __shared__ double coef[MULTI ? NC : GT_CH];
    double creg[GT_CH];
    if (MULTI) {
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) coef[j] = scal[S_VF + col0 + j];
        __syncthreads();
    } else {
#pragma unroll
        for (int j = 0; j < GT_CH; ++j) creg[j] = (j < cnt) ? scal[S_VF + col0 + j] : 0.0;
    }
    const int ncc = MULTI ? cnt : 1;
    const int64_t npair = (D + 1) >> 1;
    double2* f2 = reinterpret_cast<double2*>(f);
    const int64_t ld2 = ld >> 1;
    double acc = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < npair; r += (int64_t)gridDim.x * blockDim.x) {
        double2 fr = f2[r];
        for (int cc = 0; cc < ncc; cc += GT_CH) {
            const int nc = min(GT_CH, cnt - cc);
            const double2* V2 = reinterpret_cast<const double2*>(V + (int64_t)(col0 + cc) * ld);
            double2 v[GT_CH];
#pragma unroll
            for (int j = 0; j < GT_CH; ++j) v[j] = (j < nc) ? V2[(int64_t)j * ld2 + r] : make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < GT_CH; ++j) {
                const double c = MULTI ? ((j < nc) ? coef[cc + j] : 0.0) : creg[j];
                fr.x = fma(-v[j].x, c, fr.x);
                fr.y = fma(-v[j].y, c, fr.y);
            }
        }
        f2[r] = fr;
        acc = fma(fr.x, fr.x, fma(fr.y, fr.y, acc));
    }
    acc = bh_block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
    if (bh_last_block(counter)) {
        double t = 0.0;
        for (int bb = threadIdx.x; bb < (int)gridDim.x; bb += blockDim.x) t += part[bb];
        t = bh_block_sum(t, scratch);
        if (threadIdx.x == 0) {
            if (raw) scal[S_FLAG + 3] = t; else scal[S_BETA + i + 1] = sqrt(t);
        }
    }
}


// ---- row-partitioned solve: local raw sums are all-reduced (NCCL), then these one-thread kernels finish ----
__global__ void k_fin_sqrt(double* __restrict__ scal, int src, int dst) { scal[dst] = sqrt(scal[src]); }
__global__ void k_fin_coef(double* __restrict__ scal, int i, int pass)
{
    // classical Gram-Schmidt twice against all columns: T(i,i) and T(i,i-1) are the accumulated coefficients
    const double a = scal[S_VF + i], o = (i > 0) ? scal[S_VF + i - 1] : 0.0;
    if (pass == 0) {
        scal[S_ALPHA + i] = a;
        scal[S_OFFD + i] = o;
    } else {
        scal[S_ALPHA + i] += a;
        scal[S_OFFD + i] += o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// One Lanczos step after the H.v in ONE cooperative launch (used when the residual fits in registers:
// <= COOP_NP row pairs per thread).  The kernel keeps the residual in registers from the three-term update to
// the final normalisation:
//   w -= beta v_{i-1};  alpha = v_i.w            -> grid.sync
//   f  = w - alpha v_i
//   for each block of 8 basis columns:  c = V_B^T f -> grid.sync ;  f -= V_B c     (block MGS, V_B re-read from L2)
//   beta' = |f|                                   -> grid.sync
//   f -> memory,  v_{i+1} = f / beta'             (the next step's scaling pass is fused here)
// i.e. 2 launches per Lanczos step instead of 9-15, no f round-trips through HBM, and the basis streamed once.
// All partial sums are combined in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------------
#define COOP_NP 10
#define COOP_THREADS 256

__device__ __forceinline__ double coop_sum_partials(const double* __restrict__ part, int n, int stride)
{
    // warp-level: lanes stride over n partials spaced `stride` doubles apart, then a fixed shuffle tree
    double t = 0.0;
    for (int b = threadIdx.x & 31; b < n; b += 32) t += part[(int64_t)b * stride];
    return bh_warp_sum(t);
}

// NP = row pairs per thread (the residual lives in registers: 128 registers, two CTAs per SM).  Variants measured and
// removed in round 2: residual in shared memory with 3 CTAs per SM (-2 %), L2 prefetch of the next
// block before the barrier (-1..-6 %), skipping the update of blocks with negligible coefficients (breaks the 1e-10 parity),
// blocks of 4 columns (-3 %).
template <int CH, int NP>
__global__ void __launch_bounds__(COOP_THREADS, 2)
k_step_coop(int64_t D, int64_t ld, double* __restrict__ V, int i, int subtract, int passes, const double* __restrict__ w,
            double* __restrict__ f, double* __restrict__ scal, double* __restrict__ part, double thresh, int fused)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[COOP_THREADS / 32][CH];
    __shared__ double cs[CH];
    __shared__ double sh_val;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nblk = gridDim.x;
    const int64_t npair = (D + 1) >> 1, ld2 = ld >> 1;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x, gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double* part_a = part;                         // [nblk]
    double* part_n = part + nblk;                  // [nblk]
    double* part_c = part + 2 * (int64_t)nblk;     // [2][nblk][CH]
    const double2* w2 = reinterpret_cast<const double2*>(w);
    const double2* vi2 = reinterpret_cast<const double2*>(V + (int64_t)i * ld);
    const double2* vp2 = reinterpret_cast<const double2*>(V + (int64_t)(i > 0 ? i - 1 : 0) * ld);
    const double2* V2 = reinterpret_cast<const double2*>(V);

    // ---- three-term update and alpha ----
    const double b = subtract ? scal[S_BETA + i] : 0.0;
    double2 fr_reg[NP];
#define FR(t) fr_reg[t]
    double acc0 = 0.0;
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        const int64_t p = gtid + t * gsz;
        FR(t) = make_double2(0.0, 0.0);
        if (p < npair) {
            double2 wr = w2[p];
            if (subtract) {
                const double2 vp = vp2[p];
                wr.x = fma(-b, vp.x, wr.x);
                wr.y = fma(-b, vp.y, wr.y);
            }
            const double2 v = vi2[p];
            acc0 = fma(v.x, wr.x, fma(v.y, wr.y, acc0));
            FR(t) = wr;
        }
    }
    acc0 = bh_warp_sum(acc0);
    if (lane == 0) red[wid][0] = acc0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < COOP_THREADS / 32; ++q) t += red[q][0];
        part_a[blockIdx.x] = t;
    }
    grid.sync();
    if (wid == 0) {
        const double a = coop_sum_partials(part_a, nblk, 1);
        if (lane == 0) sh_val = a;
    }
    __syncthreads();
    double alpha = sh_val;
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        const int64_t p = gtid + t * gsz;
        if (p < npair) {
            const double2 v = vi2[p];
            FR(t).x = fma(-alpha, v.x, FR(t).x);
            FR(t).y = fma(-alpha, v.y, FR(t).y);
        }
    }

    // ---- block modified Gram-Schmidt against columns 0..i ----
    double offd = b;
    int buf = 0;
    // Fused form (default; BH_COOP_FUSED=0 selects the two-sweep form below): the update with block k (its columns re-read
    // from L2 with last-use loads, so that L2 drops them before the block being streamed in) and the dot products with block k+1
    // (streamed from HBM) run in the same sweep over the rows, so the two memory levels are busy together instead of in
    // turn; same arithmetic, same order of operations per row, one grid.sync per block as before.
    if (fused) {
        for (int pass = 0; pass < passes; ++pass) {
            double acc[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[j] = 0.0;
            {
                const int nc = min(CH, i + 1);
#pragma unroll
                for (int t = 0; t < NP; ++t) {
                    const int64_t p = gtid + t * gsz;
                    if (p < npair) {
                        double2 v[CH];
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = (j < nc) ? V2[(int64_t)j * ld2 + p] : make_double2(0.0, 0.0);
#pragma unroll
                        for (int j = 0; j < CH; ++j) acc[j] = fma(v[j].x, FR(t).x, fma(v[j].y, FR(t).y, acc[j]));
                    }
                }
            }
            for (int c0 = 0; c0 <= i; c0 += CH) {
                const int nc = min(CH, i + 1 - c0);
                // finish the dot products of block c0
#pragma unroll
                for (int j = 0; j < CH; ++j) acc[j] = bh_warp_sum(acc[j]);
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) red[wid][j] = acc[j];
                }
                __syncthreads();
                double* pc = part_c + (int64_t)buf * nblk * CH;
                if (threadIdx.x < CH) {
                    double t = 0.0;
                    for (int q = 0; q < COOP_THREADS / 32; ++q) t += red[q][threadIdx.x];
                    pc[(int64_t)blockIdx.x * CH + threadIdx.x] = t;
                }
                grid.sync();
                if (wid < nc) {
                    const double c = coop_sum_partials(pc + wid, nblk, CH);
                    if (lane == 0) cs[wid] = c;
                } else if (wid < CH && lane == 0) {
                    cs[wid] = 0.0;
                }
                __syncthreads();
                double c[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) c[j] = cs[j];
                if (i >= c0 && i < c0 + CH) alpha += cs[i - c0];                         // Lanczos.h:170
                if (subtract && i - 1 >= c0 && i - 1 < c0 + CH) offd += cs[i - 1 - c0];  // Lanczos.h:168
                // update with block c0, dot products with block c0 + CH
                const int c1 = c0 + CH;
                const int nc1 = (c1 <= i) ? min(CH, i + 1 - c1) : 0;
                const double2* VB = V2 + (int64_t)c0 * ld2;
                const double2* VN = V2 + (int64_t)c1 * ld2;
#pragma unroll
                for (int j = 0; j < CH; ++j) acc[j] = 0.0;
#pragma unroll
                for (int t = 0; t < NP; ++t) {
                    const int64_t p = gtid + t * gsz;
                    if (p < npair) {
                        double2 v[CH];
                        // last use of block c0 in this step: let L2 drop these lines first, block c0 + CH has to stay
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = (j < nc) ? __ldlu(VB + (int64_t)j * ld2 + p) : make_double2(0.0, 0.0);
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            FR(t).x = fma(-v[j].x, c[j], FR(t).x);
                            FR(t).y = fma(-v[j].y, c[j], FR(t).y);
                        }
                        if (nc1 > 0) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) v[j] = (j < nc1) ? VN[(int64_t)j * ld2 + p] : make_double2(0.0, 0.0);
#pragma unroll
                            for (int j = 0; j < CH; ++j) acc[j] = fma(v[j].x, FR(t).x, fma(v[j].y, FR(t).y, acc[j]));
                        }
                    }
                }
                buf ^= 1;
                __syncthreads();  // cs / red are rewritten by the next block of columns
            }
        }
    } else
    for (int pass = 0; pass < passes; ++pass) {
        for (int c0 = 0; c0 <= i; c0 += CH) {
            const int nc = min(CH, i + 1 - c0);
            const double2* VB = V2 + (int64_t)c0 * ld2;
            double acc[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[j] = 0.0;
#pragma unroll
            for (int t = 0; t < NP; ++t) {
                const int64_t p = gtid + t * gsz;
                if (p < npair) {
                    double2 v[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = (j < nc) ? VB[(int64_t)j * ld2 + p] : make_double2(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < CH; ++j) acc[j] = fma(v[j].x, FR(t).x, fma(v[j].y, FR(t).y, acc[j]));
                }
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[j] = bh_warp_sum(acc[j]);
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < CH; ++j) red[wid][j] = acc[j];
            }
            __syncthreads();
            double* pc = part_c + (int64_t)buf * nblk * CH;
            if (threadIdx.x < CH) {
                double t = 0.0;
                for (int q = 0; q < COOP_THREADS / 32; ++q) t += red[q][threadIdx.x];
                pc[(int64_t)blockIdx.x * CH + threadIdx.x] = t;
            }
            grid.sync();
            if (wid < nc) {
                const double c = coop_sum_partials(pc + wid, nblk, CH);
                if (lane == 0) cs[wid] = c;
            } else if (wid < CH && lane == 0) {
                cs[wid] = 0.0;
            }
            __syncthreads();
            double c[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) c[j] = cs[j];
            if (i >= c0 && i < c0 + CH) alpha += cs[i - c0];                         // Lanczos.h:170
            if (subtract && i - 1 >= c0 && i - 1 < c0 + CH) offd += cs[i - 1 - c0];  // Lanczos.h:168
#pragma unroll
            for (int t = 0; t < NP; ++t) {
                const int64_t p = gtid + t * gsz;
                if (p < npair) {
                    double2 v[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = (j < nc) ? VB[(int64_t)j * ld2 + p] : make_double2(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        FR(t).x = fma(-v[j].x, c[j], FR(t).x);
                        FR(t).y = fma(-v[j].y, c[j], FR(t).y);
                    }
                }
            }
            buf ^= 1;
            __syncthreads();  // cs / red are rewritten by the next block of columns
        }
    }

    // ---- norm, write-back, next basis vector ----
    double nrm = 0.0;
#pragma unroll
    for (int t = 0; t < NP; ++t) nrm = fma(FR(t).x, FR(t).x, fma(FR(t).y, FR(t).y, nrm));
    nrm = bh_warp_sum(nrm);
    if (lane == 0) red[wid][0] = nrm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < COOP_THREADS / 32; ++q) t += red[q][0];
        part_n[blockIdx.x] = t;
    }
    grid.sync();
    if (wid == 0) {
        const double t = coop_sum_partials(part_n, nblk, 1);
        if (lane == 0) sh_val = sqrt(t);
    }
    __syncthreads();
    const double beta = sh_val;
    const double inv = (beta > thresh) ? 1.0 / beta : 0.0;
    double2* f2 = reinterpret_cast<double2*>(f);
    double2* vn2 = reinterpret_cast<double2*>(V + (int64_t)(i + 1) * ld);
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        const int64_t p = gtid + t * gsz;
        if (p < npair) {
            f2[p] = FR(t);
            vn2[p] = make_double2(FR(t).x * inv, FR(t).y * inv);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[S_ALPHA + i] = alpha;
        scal[S_OFFD + i] = offd;
        scal[S_BETA + i + 1] = beta;
        if (!(beta > thresh) && scal[S_FLAG] == 0.0) {
            scal[S_FLAG] = 1.0;
            scal[S_FLAG + 1] = (double)(i + 1);
        }
    }
}
#undef FR

// x *= 1 / scal[idx]  (final normalisation of a Ritz vector: under partial re-orthogonalisation the basis is
// orthonormal only to sqrt(eps), so |V y| differs from 1 at that level)
__global__ void __launch_bounds__(VEC_THREADS)
k_scale_inplace(int64_t D, double* __restrict__ x, const double* __restrict__ scal, int idx)
{
    const double inv = 1.0 / scal[idx];
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < D; r += (int64_t)gridDim.x * blockDim.x) x[r] *= inv;
}

// out = V[:, 0..cnt) y   (Ritz vector, HermEigsBase.h:456-479)
__global__ void __launch_bounds__(VEC_THREADS)
k_lincomb(int64_t D, int64_t ld, const double* __restrict__ V, int cnt, const double* __restrict__ y,
          double* __restrict__ out)
{
    __shared__ double coef[NC];
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) coef[j] = y[j];
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < D; r += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int j = 0; j < cnt; ++j) acc += V[(int64_t)j * ld + r] * coef[j];
        out[r] = acc;
    }
}

// K6: V[:, 0..k) <- V[:, 0..ncv) Y[:, 0..k), in place (a CTA owns CR rows: it reads all their ncv entries
// into shared memory before it writes any of them).  Y is ncv x k column-major in global memory.
#define CK 16
template <int CR>
__global__ void __launch_bounds__(256)
k_compress(int64_t D, int64_t ld, double* __restrict__ V, int ncv, int k, const double* __restrict__ Y)
{
    extern __shared__ double sm[];
    double* vt = sm;                      // [ncv][CR]
    double* yc = sm + (size_t)ncv * CR;   // [CK][ncv]
    const int64_t r0 = (int64_t)blockIdx.x * CR;
    const int rr = threadIdx.x % CR, cg = threadIdx.x / CR;  // 256 / CR column groups
    for (int idx = threadIdx.x; idx < ncv * CR; idx += blockDim.x) {
        const int j = idx / CR, r = idx % CR;
        vt[idx] = (r0 + r < D) ? V[(int64_t)j * ld + r0 + r] : 0.0;
    }
    for (int c0 = 0; c0 < k; c0 += CK) {
        const int nck = min(CK, k - c0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < nck * ncv; idx += blockDim.x) yc[idx] = Y[(size_t)c0 * ncv + idx];
        __syncthreads();
        for (int l = cg; l < nck; l += 256 / CR) {
            const double* yl = yc + (size_t)l * ncv;
            double acc = 0.0;
            for (int j = 0; j < ncv; ++j) acc += vt[j * CR + rr] * yl[j];
            if (r0 + rr < D) V[(int64_t)(c0 + l) * ld + r0 + rr] = acc;
        }
    }
}

// K6, register-tiled: a CTA owns 128 rows (staged once in shared memory, so the update is safe in place), every
// thread accumulates a 4 x 4 tile (4 rows x 4 output columns) -> 16 FMAs per 4 shared-memory loads instead of 1 per 2.
#define CT_ROWS 128
#define CT_COLS 32
__global__ void __launch_bounds__(256)
k_compress_tiled(int64_t D, int64_t ld, double* __restrict__ V, int ncv, int k, const double* __restrict__ Y)
{
    extern __shared__ __align__(16) double sm2[];
    double* vt = sm2;                           // [ncv][CT_ROWS]
    double* yc = sm2 + (size_t)ncv * CT_ROWS;   // [ncv][CT_COLS]   (row j = coefficients of basis vector j)
    const int64_t r0 = (int64_t)blockIdx.x * CT_ROWS;
    const int rg = threadIdx.x & 31, cgp = threadIdx.x >> 5;  // 32 row groups x 8 column groups
    for (int idx = threadIdx.x; idx < ncv * CT_ROWS; idx += blockDim.x) {
        const int j = idx / CT_ROWS, r = idx % CT_ROWS;
        vt[idx] = (r0 + r < D) ? V[(int64_t)j * ld + r0 + r] : 0.0;
    }
    for (int c0 = 0; c0 < k; c0 += CT_COLS) {
        const int nck = min(CT_COLS, k - c0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < ncv * CT_COLS; idx += blockDim.x) {
            const int j = idx / CT_COLS, c = idx % CT_COLS;
            yc[idx] = (c < nck) ? Y[(size_t)(c0 + c) * ncv + j] : 0.0;
        }
        __syncthreads();
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        for (int j = 0; j < ncv; ++j) {
            const double2 a01 = *reinterpret_cast<const double2*>(vt + j * CT_ROWS + rg * 4);
            const double2 a23 = *reinterpret_cast<const double2*>(vt + j * CT_ROWS + rg * 4 + 2);
            const double2 b01 = *reinterpret_cast<const double2*>(yc + j * CT_COLS + cgp * 4);
            const double2 b23 = *reinterpret_cast<const double2*>(yc + j * CT_COLS + cgp * 4 + 2);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y};
            const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int ra = 0; ra < 4; ++ra)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) acc[ra][cb] = fma(a[ra], b[cb], acc[ra][cb]);
        }
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const int c = cgp * 4 + cb;
            if (c < nck) {
                double* dst = V + (int64_t)(c0 + c) * ld + r0 + rg;
#pragma unroll
                for (int ra = 0; ra < 4; ++ra)
                    if (r0 + rg + 64 * ra < D) dst[64 * ra] = acc[ra][cb];
            }
        }
    }
}

// K6, wider register tile: a thread owns 4 rows x 8 columns (32 accumulators), a CTA 256 rows x 32 columns per sweep.
// Per j the warp reads 8 shared-memory wavefronts of V (4 consecutive doubles per lane) and 2 broadcast wavefronts of Y for
// 1024 FMAs (16 issue cycles of the FP64 pipe): bound by the FP64 rate instead of by shared memory like the 4 x 4 tile.
#define CT8_ROWS 256
__global__ void __launch_bounds__(256, 2)
k_compress_tiled8(int64_t D, int64_t ld, double* __restrict__ V, int ncv, int k, const double* __restrict__ Y)
{
    extern __shared__ __align__(16) double sm3[];
    double* vt = sm3;                            // [ncv][CT8_ROWS]
    double* yc = sm3 + (size_t)ncv * CT8_ROWS;   // [ncv][CT_COLS]
    const int64_t r0 = (int64_t)blockIdx.x * CT8_ROWS;
    const int rg = threadIdx.x & 63, cgp = threadIdx.x >> 6;  // 64 row groups (rows rg, rg + 64, rg + 128, rg + 192) x 4 column groups (8 columns)
    for (int idx = threadIdx.x; idx < ncv * CT8_ROWS; idx += blockDim.x) {
        const int j = idx / CT8_ROWS, r = idx % CT8_ROWS;
        vt[idx] = (r0 + r < D) ? V[(int64_t)j * ld + r0 + r] : 0.0;
    }
    for (int c0 = 0; c0 < k; c0 += CT_COLS) {
        const int nck = min(CT_COLS, k - c0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < ncv * CT_COLS; idx += blockDim.x) {
            const int j = idx / CT_COLS, c = idx % CT_COLS;
            yc[idx] = (c < nck) ? Y[(size_t)(c0 + c) * ncv + j] : 0.0;
        }
        __syncthreads();
        double acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
        for (int j = 0; j < ncv; ++j) {
            // the 4 rows of a thread are 64 apart: consecutive lanes read consecutive doubles (conflict-free; 4 consecutive rows per
            // thread made every 128-bit load a 2-way bank conflict: 14.5 M conflicts per launch under ncu) and the stores coalesce
            double a[4];
#pragma unroll
            for (int ra = 0; ra < 4; ++ra) a[ra] = vt[j * CT8_ROWS + rg + 64 * ra];
            double b[8];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const double2 t = *reinterpret_cast<const double2*>(yc + j * CT_COLS + cgp * 8 + 2 * h);
                b[2 * h] = t.x;
                b[2 * h + 1] = t.y;
            }
#pragma unroll
            for (int ra = 0; ra < 4; ++ra)
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) acc[ra][cb] = fma(a[ra], b[cb], acc[ra][cb]);
        }
#pragma unroll
        for (int cb = 0; cb < 8; ++cb) {
            const int c = cgp * 8 + cb;
            if (c < nck) {
                double* dst = V + (int64_t)(c0 + c) * ld + r0 + rg;
#pragma unroll
                for (int ra = 0; ra < 4; ++ra)
                    if (r0 + rg + 64 * ra < D) dst[64 * ra] = acc[ra][cb];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Gram product M = V^T W (ncv x ncv, ncv <= GRAM_N) for the Rayleigh-Ritz step of the accelerated solver: one pass
// over V and W instead of ncv transposed products that each re-read V.  A CTA stages GRAM_R rows of both matrices
// in shared memory ([row][column], padded), every thread keeps a 3 x 3 tile of M in registers across all its row
// tiles; per-CTA partials are combined in a fixed order by k_gram_reduce (deterministic).
// ---------------------------------------------------------------------------------------------------------
#define GRAM_N 48
#define GRAM_R 32
#define GRAM_LD 49
__global__ void __launch_bounds__(256, 2)
k_gram(int64_t D, int64_t ld, const double* __restrict__ V, const double* __restrict__ W, int ncv, double* __restrict__ part)
{
    __shared__ double vt[GRAM_R * GRAM_LD];
    __shared__ double wt[GRAM_R * GRAM_LD];
    const int ta = threadIdx.x >> 4, tb = threadIdx.x & 15;  // tile (3 ta .. 3 ta + 2) x (3 tb .. 3 tb + 2)
    double acc[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[a][b] = 0.0;
    for (int idx = threadIdx.x; idx < GRAM_R * GRAM_LD; idx += blockDim.x) {
        vt[idx] = 0.0;
        wt[idx] = 0.0;
    }
    const int64_t ntiles = (D + GRAM_R - 1) / GRAM_R;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * GRAM_R;
        __syncthreads();
        for (int idx = threadIdx.x; idx < ncv * GRAM_R; idx += blockDim.x) {
            const int j = idx / GRAM_R, r = idx % GRAM_R;
            const bool in = r0 + r < D;
            vt[r * GRAM_LD + j] = in ? V[(int64_t)j * ld + r0 + r] : 0.0;
            wt[r * GRAM_LD + j] = in ? W[(int64_t)j * ld + r0 + r] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < GRAM_R; ++r) {
            double a[3], b[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                a[q] = vt[r * GRAM_LD + 3 * ta + q];
                b[q] = wt[r * GRAM_LD + 3 * tb + q];
            }
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int y = 0; y < 3; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
    }
    double* dst = part + (size_t)blockIdx.x * GRAM_N * GRAM_N;
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) dst[(3 * ta + x) * GRAM_N + 3 * tb + y] = acc[x][y];
}

// M[a + b * ncv] = sum over CTAs of part[cta][a][b], fixed order
__global__ void k_gram_reduce(int nparts, const double* __restrict__ part, int ncv, double* __restrict__ M)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncv * ncv) return;
    const int a = idx % ncv, b = idx / ncv;
    double t = 0.0;
    for (int p = 0; p < nparts; ++p) t += part[(size_t)p * GRAM_N * GRAM_N + a * GRAM_N + b];
    M[a + (size_t)b * ncv] = t;
}

int bh_ensure_workspace(bh_ctx* ctx, int ncv)
{
    if (!ctx->D) return bh_fail(ctx, BH_ERR_STATE, "no system: call bh_setup first");
    // row-partitioned chain context: basis, w, f and the Chebyshev buffers live in one arena that the other ranks map (dist.cu)
    if (bh_dist_peer_wanted(ctx)) BH_TRY(bh_dist_arena(ctx, std::max(ncv, 12)));
    // vectors are padded to ld (a multiple of 32) with zeros: the re-orthogonalisation kernels read row pairs
    if (!ctx->d_w) {
        BH_CUDA(ctx, cudaMalloc(&ctx->d_w, sizeof(double) * ctx->ld));
        BH_CUDA(ctx, cudaMemsetAsync(ctx->d_w, 0, sizeof(double) * ctx->ld, ctx->stream));
    }
    if (!ctx->d_f) {
        BH_CUDA(ctx, cudaMalloc(&ctx->d_f, sizeof(double) * ctx->ld));
        BH_CUDA(ctx, cudaMemsetAsync(ctx->d_f, 0, sizeof(double) * ctx->ld, ctx->stream));
    }
    if (!ctx->d_scal) {
        BH_CUDA(ctx, cudaMalloc(&ctx->d_scal, sizeof(double) * S_TOTAL));
        BH_CUDA(ctx, cudaMalloc(&ctx->d_part, sizeof(double) * (size_t)NC * (ctx->sm_count * 8)));
        BH_CUDA(ctx, cudaMalloc(&ctx->d_counter, sizeof(unsigned int) * 4));
        BH_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int) * 4, ctx->stream));
        BH_CUDA(ctx, cudaMalloc(&ctx->d_small, sizeof(double) * (size_t)NC * NC));
    }
    if (ncv > ctx->ws_ncv) {
        if (ctx->d_V) cudaFree(ctx->d_V);
        ctx->d_V = nullptr;
        BH_CUDA(ctx, cudaMalloc(&ctx->d_V, sizeof(double) * (size_t)ctx->ld * (ncv + 1)));
        BH_CUDA(ctx, cudaMemsetAsync(ctx->d_V, 0, sizeof(double) * (size_t)ctx->ld * (ncv + 1), ctx->stream));
        ctx->ws_ncv = ncv;
    }
    return BH_OK;
}

// y = Op x on device vectors (local slices); returns a BH_* status
typedef std::function<int(const double*, double*)> LanczosOp;

// Thick-restart Lanczos for the nev algebraically smallest eigenvalues of `op`.  start_given: the start vector is
// already in ctx->d_w (otherwise Spectra's LCG vector is used); like Spectra it is pre-multiplied by the operator.
static int lanczos_core(bh_ctx* ctx, const LanczosOp& op, bool start_given, int nev, int ncv, double tol, int maxit,
                        BhSolve* out)
{
    const int64_t Dglobal = ctx->D, ld = ctx->ld;
    const int64_t D = ctx->nloc;  // rows held by this context (all of them unless the solve is row-partitioned)
    const bool dist = ctx->partitioned && ctx->world > 1;
    // Spectra::GenEigsBase constructor checks (GenEigsBase.h:335-339), which the reference relies on
    if (nev < 1 || nev > Dglobal - 2) return bh_fail(ctx, BH_ERR_ARG, "nev must satisfy 1 <= nev <= n - 2, n is the size of matrix");
    if (ncv < nev + 2 || ncv > Dglobal) return bh_fail(ctx, BH_ERR_ARG, "ncv must satisfy nev + 2 <= ncv <= n, n is the size of matrix");
    if (ncv > BH_MAX_NCV) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "ncv exceeds BH_MAX_NCV");
    const auto t_start = std::chrono::steady_clock::now();
    BH_TRY(bh_ensure_workspace(ctx, ncv));
    cudaStream_t st = ctx->stream;
    double* V = ctx->d_V;
    double* scal = ctx->d_scal;
    double* part = ctx->d_part;
    unsigned int* counter = ctx->d_counter;
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks(D, VEC_THREADS), (int64_t)ctx->sm_count * 4));
    const double eps = std::numeric_limits<double>::epsilon();
    const double near0 = std::numeric_limits<double>::min() * 10.0;
    const double eps23 = std::pow(eps, 2.0 / 3.0);

    BH_CUDA(ctx, cudaMemsetAsync(scal, 0, sizeof(double) * S_TOTAL, st));
    // v0 = LCG ; f = A v0 ; beta[0] = |f|   (Arnoldi.h:147-154)
    if (!start_given) BH_TRY(bh_lcg_fill_dev(ctx, ctx->d_w, D));
    BH_TRY(op(ctx->d_w, ctx->d_f));
    if (dist) {
        k_norm<<<G, VEC_THREADS, 0, st>>>(D, ctx->d_f, scal, S_FLAG + 3, part, counter, 1);
        BH_TRY(bh_dist_allreduce_sum(ctx, scal + S_FLAG + 3, 1));
        k_fin_sqrt<<<1, 1, 0, st>>>(scal, S_FLAG + 3, S_BETA + 0);
    } else {
        k_norm<<<G, VEC_THREADS, 0, st>>>(D, ctx->d_f, scal, S_BETA + 0, part, counter);
    }
    BH_LAUNCHED(ctx);
    int nmatvec = 1;

    std::vector<double> h_scal(S_TOTAL);
    std::vector<double> theta, coup;  // kept Ritz values and their coupling to the next vector
    std::vector<double> T, evals, Y;
    int from = 0, nconv = 0, iter = 0;
    double beta_last = 0.0;
    const int CRsel = (ncv <= 160) ? 64 : 16;
    // column block of the re-orthogonalisation: L2-sized blocks for big systems, everything at once when the
    // whole basis is L2-resident anyway (then launch count is what matters)
    const int rblock = ((size_t)ld * 8 * (ncv + 1) <= ((size_t)48 << 20) && !ctx->reorth_block_forced) ? ncv + 1 : ctx->reorth_block;
    // The opt-in shared-memory limits are properties of the FUNCTIONS (shared by every context and lockstep solve on the
    // device): they are always raised to the fixed maxima the selection rules below allow, never to a per-call size -- solves
    // with different ncv run interleaved (lockstep batches; the quick stage 1 uses a shorter basis than stage 2).
    const size_t tiled_smem = sizeof(double) * (size_t)ncv * (CT_ROWS + CT_COLS);
    const bool tiled = tiled_smem <= 200 * 1024 && ctx->compress_tiled;
    if (tiled) BH_CUDA(ctx, cudaFuncSetAttribute(k_compress_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t tiled8_smem = sizeof(double) * (size_t)ncv * (CT8_ROWS + CT_COLS);
    const bool tiled8 = tiled && ctx->compress_tiled >= 2 && tiled8_smem <= 110 * 1024;
    if (tiled8) BH_CUDA(ctx, cudaFuncSetAttribute(k_compress_tiled8, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    const size_t compress_smem = sizeof(double) * ((size_t)ncv * CRsel + (size_t)CK * ncv);
    if (compress_smem > 227 * 1024) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "restart kernel: ncv too large for shared memory");
    if (compress_smem > 48 * 1024) {
        if (CRsel == 64)
            BH_CUDA(ctx, cudaFuncSetAttribute(k_compress<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else
            BH_CUDA(ctx, cudaFuncSetAttribute(k_compress<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }

    // cooperative single-launch step: needs every CTA resident and <= COOP_NP row pairs per thread
    int coop_grid = 0;
    if (ctx->coop && !dist) {
        int bps = 0;
        const int64_t npair = (D + 1) / 2;
        BH_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_step_coop<GT_CH, COOP_NP>, COOP_THREADS, 0));
        bps = std::min(bps, 2);
        if (bps >= 1 && npair <= (int64_t)ctx->sm_count * bps * COOP_THREADS * COOP_NP) {
            coop_grid = ctx->sm_count * bps;
            // do not spread a tiny problem over idle CTAs: grid syncs cost more with more CTAs
            const int64_t need = (npair + COOP_THREADS - 1) / COOP_THREADS;
            if (need < coop_grid) coop_grid = (int)std::max<int64_t>(need, 1);
        }
    }

    // steps that start from a fresh random vector after a breakdown (Spectra's expand_basis, Arnoldi.h:64-113):
    // no coupling to the previous column, re-orthogonalised twice
    std::vector<char> fresh(ncv + 2, 0);
    int istart = 0, nfresh = 0;
    for (;;) {
        for (int i = istart; i < ncv; ++i) {
            double* vi = V + (int64_t)i * ld;
            const bool first_after_restart = (from > 0 && i == from) || fresh[i];
            // v_i = f / beta: written by the previous cooperative step except at the start of a cycle
            if (!coop_grid || i == istart) {
                k_scale<<<G, VEC_THREADS, 0, st>>>(D, ctx->d_f, vi, scal, i, near0);
                BH_LAUNCHED(ctx);
            }
            BH_TRY(op(vi, ctx->d_w));
            ++nmatvec;
            const int subtract = (i > 0 && !first_after_restart) ? 1 : 0;
            if (dist) {
                // classical Gram-Schmidt twice against every column (it subsumes the three-term recurrence and the
                // coupling row after a restart); 4 small all-reduces per step
                std::swap(ctx->d_w, ctx->d_f);  // the residual is built in place in what H.v just wrote
                for (int pass = 0; pass < 2; ++pass) {
                    k_gemv_t<true><<<G, VEC_THREADS, 0, st>>>(D, ld, V, 0, i + 1, ctx->d_f, scal, i, 0, part, counter, 1);
                    BH_TRY(bh_dist_allreduce_sum(ctx, scal + S_VF, i + 1));
                    k_fin_coef<<<1, 1, 0, st>>>(scal, i, pass);
                    k_gemv_n_norm<true><<<G, VEC_THREADS, 0, st>>>(D, ld, V, 0, i + 1, ctx->d_f, scal, i, part, counter, 1);
                    if (pass == 1) BH_TRY(bh_dist_allreduce_sum(ctx, scal + S_FLAG + 3, 1));  // only the final norm is used
                }
                k_fin_sqrt<<<1, 1, 0, st>>>(scal, S_FLAG + 3, S_BETA + i + 1);
                ctx->launches += 8;
                continue;
            }
            if (coop_grid) {
                int ii = i, sub = subtract, npass = first_after_restart ? 2 : 1;
                int64_t Dv = D, ldv = ld;
                const double* wv = ctx->d_w;
                double* fv = ctx->d_f;
                double th = near0;
                int fused = ctx->coop_fused;
                void* args[] = {&Dv, &ldv, &V, &ii, &sub, &npass, &wv, &fv, &scal, &part, &th, &fused};
                void* fn = (void*)k_step_coop<GT_CH, COOP_NP>;
                const size_t fr_bytes = 0;
                {
                    // algorithmic bytes: w and the i + 1 basis columns read once, f and v_{i+1} written
                    BhProfScope prof(ctx, BH_PROF_STEP, 8.0 * (double)D * (i + 1 + 3));
                    BH_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(coop_grid), dim3(COOP_THREADS), args, fr_bytes, st));
                }
                BH_LAUNCHED(ctx);
                continue;
            }
            BhProfScope prof(ctx, BH_PROF_STEP, 8.0 * (double)D * (i + 1 + 3));
            k_local_alpha<<<G, VEC_THREADS, 0, st>>>(D, ctx->d_w, i > 0 ? vi - ld : vi, vi, scal, i, subtract, part, counter);
            k_resid_norm<<<G, VEC_THREADS, 0, st>>>(D, ctx->d_w, vi, ctx->d_f, scal, i, part, counter);
            const int passes = first_after_restart ? 2 : 1;
            int nlaunch = 2;
            for (int p = 0; p < passes; ++p) {
                for (int c0 = 0; c0 <= i; c0 += rblock) {
                    const int cnt = std::min(rblock, i + 1 - c0);
                    if (rblock <= GT_CH) {
                        k_gemv_t<false><<<G, VEC_THREADS, 0, st>>>(D, ld, V, c0, cnt, ctx->d_f, scal, i, subtract, part, counter);
                        k_gemv_n_norm<false><<<G, VEC_THREADS, 0, st>>>(D, ld, V, c0, cnt, ctx->d_f, scal, i, part, counter);
                    } else {
                        k_gemv_t<true><<<G, VEC_THREADS, 0, st>>>(D, ld, V, c0, cnt, ctx->d_f, scal, i, subtract, part, counter);
                        k_gemv_n_norm<true><<<G, VEC_THREADS, 0, st>>>(D, ld, V, c0, cnt, ctx->d_f, scal, i, part, counter);
                    }
                    nlaunch += 2;
                }
            }
            ctx->launches += nlaunch;
        }
        BH_CUDA(ctx, cudaGetLastError());
        BH_D2H(ctx, h_scal.data(), scal, sizeof(double) * S_TOTAL);
        BH_CUDA(ctx, cudaStreamSynchronize(st));
        if (h_scal[S_FLAG] != 0.0) {
            // invariant subspace before the basis was full (e.g. J = 0: H diagonal with few distinct levels).
            // Like Spectra: continue from a new random vector orthogonal to the columns built so far.
            const int ib = (int)h_scal[S_FLAG + 1];
            if (++nfresh > 4 * ncv || ib < 0 || ib >= ncv)
                return bh_fail(ctx, BH_ERR_NOCONV, "Lanczos breakdown (invariant subspace reached before ncv steps)");
            BH_CUDA(ctx, cudaMemsetAsync(scal + S_FLAG, 0, sizeof(double) * 2, st));
            const int64_t first = (ctx->partitioned ? ctx->row0 : 0) + (int64_t)7919 * nfresh;
            k_lcg_seq<<<nblocks((D + 63) / 64, 128), 128, 0, st>>>(first, D, ctx->d_f);
            BH_LAUNCHED(ctx);
            if (ib > 0) {
                for (int pass = 0; pass < 2; ++pass) {
                    k_gemv_t<true><<<G, VEC_THREADS, 0, st>>>(D, ld, V, 0, ib, ctx->d_f, scal, ib - 1, 0, part, counter, 1);
                    if (dist) BH_TRY(bh_dist_allreduce_sum(ctx, scal + S_VF, ib));
                    k_gemv_n_norm<true><<<G, VEC_THREADS, 0, st>>>(D, ld, V, 0, ib, ctx->d_f, scal, ib - 1, part, counter, dist ? 1 : 0);
                    if (dist) {
                        BH_TRY(bh_dist_allreduce_sum(ctx, scal + S_FLAG + 3, 1));
                        k_fin_sqrt<<<1, 1, 0, st>>>(scal, S_FLAG + 3, S_BETA + ib);
                    }
                }
            } else if (dist) {
                k_norm<<<G, VEC_THREADS, 0, st>>>(D, ctx->d_f, scal, S_FLAG + 3, part, counter, 1);
                BH_TRY(bh_dist_allreduce_sum(ctx, scal + S_FLAG + 3, 1));
                k_fin_sqrt<<<1, 1, 0, st>>>(scal, S_FLAG + 3, S_BETA + 0);
            } else {
                k_norm<<<G, VEC_THREADS, 0, st>>>(D, ctx->d_f, scal, S_BETA + 0, part, counter);
            }
            fresh[ib] = 1;
            istart = ib;
            continue;
        }
        // projected matrix: diag(theta) + coupling row at `from`, tridiagonal afterwards
        T.assign((size_t)ncv * ncv, 0.0);
        for (int l = 0; l < from; ++l) {
            T[l + (size_t)l * ncv] = theta[l];
            T[from + (size_t)l * ncv] = T[l + (size_t)from * ncv] = coup[l];
        }
        for (int i = from; i < ncv; ++i) {
            T[i + (size_t)i * ncv] = h_scal[S_ALPHA + i];
            if (i > from) T[i + (size_t)(i - 1) * ncv] = T[(i - 1) + (size_t)i * ncv] = h_scal[S_OFFD + i];
        }
        bh_sym_eig(ncv, T, evals, Y);  // ascending = Spectra's SmallestAlge selection
        beta_last = h_scal[S_BETA + ncv];
        nconv = 0;
        for (int l = 0; l < nev; ++l) {
            const double thresh = tol * std::max(eps23, std::fabs(evals[l]));
            const double resid = std::fabs(Y[(ncv - 1) + (size_t)l * ncv]) * beta_last;
            nconv += (resid < thresh);
        }
        if (nconv >= nev || iter >= maxit) break;
        // quick-mode filter with a misplaced cut (tools/model_cut.py): a healthy run has the nev-th Ritz value below -1 after the
        // FIRST cycle, a cut below E_{nev-1} keeps it inside the damped band for ever -- stop after a few cycles instead of 80
        // (>= 80 filtered steps as well: with a short basis, e.g. nev = 2 / ncv = 12, three cycles are too few to judge)
        if (ctx->cheb_stall_iter > 0 && iter >= ctx->cheb_stall_iter && nmatvec >= 80 && evals[nev - 1] > -1.0) break;
        ++iter;
        // HermEigsBase.h:172-196
        int knew = nev;
        for (int l = nev; l < ncv; ++l)
            if (std::fabs(Y[(ncv - 1) + (size_t)l * ncv]) < near0) ++knew;
        knew += std::min(nconv, (ncv - knew) / 2);
        if (knew == 1 && ncv >= 6)
            knew = ncv / 2;
        else if (knew == 1 && ncv > 2)
            knew = 2;
        if (knew > ncv - 1) knew = ncv - 1;
        // K6: V[:, 0..knew) <- V Y[:, 0..knew)
        BH_H2D(ctx, ctx->d_small, Y.data(), sizeof(double) * (size_t)ncv * knew);
        BhProfScope prof(ctx, BH_PROF_RESTART, 8.0 * (double)D * (ncv + knew));
        if (tiled8)
            k_compress_tiled8<<<nblocks(D, CT8_ROWS), 256, tiled8_smem, st>>>(D, ld, V, ncv, knew, ctx->d_small);
        else if (tiled)
            k_compress_tiled<<<nblocks(D, CT_ROWS), 256, tiled_smem, st>>>(D, ld, V, ncv, knew, ctx->d_small);
        else if (CRsel == 64)
            k_compress<64><<<nblocks(D, 64), 256, compress_smem, st>>>(D, ld, V, ncv, knew, ctx->d_small);
        else
            k_compress<16><<<nblocks(D, 16), 256, compress_smem, st>>>(D, ld, V, ncv, knew, ctx->d_small);
        BH_LAUNCHED(ctx);
        theta.assign(evals.begin(), evals.begin() + knew);
        coup.resize(knew);
        for (int l = 0; l < knew; ++l) coup[l] = beta_last * Y[(ncv - 1) + (size_t)l * ncv];
        // the residual f and its norm carry over: V_knew = f / beta_last
        BH_CUDA(ctx, cudaMemcpyAsync(scal + S_BETA + knew, scal + S_BETA + ncv, sizeof(double), cudaMemcpyDeviceToDevice, st));
        from = knew;
        istart = from;
        std::fill(fresh.begin(), fresh.end(), 0);
    }
    out->nev = nev;
    out->ncv = ncv;
    out->evals.assign(evals.begin(), evals.begin() + nev);
    out->Y.assign(Y.begin(), Y.begin() + (size_t)ncv * nev);
    out->info.nconv = std::min(nconv, nev);
    out->info.nmatvec = nmatvec;
    out->info.nrestart = iter + 1;
    out->info.nreorth = nmatvec - 1;  // every step re-orthogonalises (see DESIGN.md on partial re-orthogonalisation)
    out->info.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    if (nconv < nev) return bh_fail(ctx, BH_ERR_NOCONV, "Eigenvalue computation failed.");
    return BH_OK;
}

// One attempt of the accelerated solve (see bh_lanczos).  *retry = true: the quick stage 1 misplaced the cut, run the full form.
static int accel_solve(bh_ctx* ctx, const LanczosOp& plain, int& hv_count, double cJ, double cU, double cmu, int nev, int ncv, double tol,
                       int maxit, int kernel, bool quick, bool* retry, BhSolve* out)
{
    const int d = ctx->cheb_degree;
    *retry = false;
    // ---- stage 1 ----
    BhSolve s1;
    int rc;
    if (quick) {
        rc = lanczos_core(ctx, plain, false, 1, ctx->cheb_quick, tol, 0, &s1);
        if (rc != BH_ERR_NOCONV && rc != BH_OK) return rc;
    } else {
        rc = lanczos_core(ctx, plain, false, nev, ncv, tol, std::min(maxit, ctx->cheb_pre), &s1);
        if (rc != BH_ERR_NOCONV || (int)s1.evals.size() < nev) {  // converged already, or a real error
            *out = s1;
            out->info.nmatvec = hv_count;
            return rc;
        }
    }
    double lo = 0, hi = 0;
    BH_TRY(bh_spectrum_bounds(ctx, cJ, cU, cmu, &lo, &hi));
    const double th0 = s1.evals[0], thn = quick ? s1.evals[0] : s1.evals[nev - 1];
    hi += 1e-9 * (hi - lo) + 1e-12;
    double cut = std::max(thn + ctx->cheb_margin * (thn - th0), th0 + ctx->cheb_frac * (hi - th0));
    if (!(cut < th0 + 0.8 * (hi - th0))) {  // no room for a filter: finish with the plain solver
        if (quick) { *retry = true; return BH_OK; }
        rc = lanczos_core(ctx, plain, false, nev, ncv, tol, maxit, out);
        out->info.nmatvec = hv_count;
        return rc;
    }
    const double c = 0.5 * (hi + cut), e = 0.5 * (hi - cut);
    // start vector of stage 2: the sum of the stage-1 Ritz vectors of the wanted end (quick: the ground Ritz vector)
    {
        const int n1 = s1.ncv, take = quick ? 1 : nev;
        std::vector<double> ysum(n1, 0.0);
        for (int l = 0; l < take; ++l)
            for (int r = 0; r < n1; ++r) ysum[r] += s1.Y[r + (size_t)l * n1];
        BH_H2D(ctx, ctx->d_small, ysum.data(), sizeof(double) * n1);
        const int G = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks(ctx->nloc, VEC_THREADS), (int64_t)ctx->sm_count * 4));
        k_lincomb<<<G, VEC_THREADS, 0, ctx->stream>>>(ctx->nloc, ctx->ld, ctx->d_V, n1, ctx->d_small, ctx->d_w);
        BH_LAUNCHED(ctx);
    }
    for (int q = 0; q < 3; ++q)
        if (!ctx->d_cheb[q]) {
            BH_CUDA(ctx, cudaMalloc(&ctx->d_cheb[q], sizeof(double) * ctx->ld));
            BH_CUDA(ctx, cudaMemsetAsync(ctx->d_cheb[q], 0, sizeof(double) * ctx->ld, ctx->stream));
        }
    // ---- stage 2 ----
    LanczosOp cheb = [&](const double* x, double* y) {
        if (ctx->parent && ctx->parent->hub) {  // lockstep solve (batch.cu): the filters of all live solves share their launches
            bool handled = false;
            BH_TRY(bh_batch_filter(ctx, x, y, c, e, cJ, cU, cmu, d, &handled));
            if (handled) {
                hv_count += d;
                return (int)BH_OK;
            }
        }
        const double* tkm2 = x;        // T_{k-2}
        const double* tkm1 = nullptr;  // T_{k-1}
        for (int k = 1; k <= d; ++k) {
            BhEpilogue ep;
            if (k == 1) {
                ep.s1 = 1.0 / e; ep.s2 = -c / e;
            } else {
                ep.s1 = 2.0 / e; ep.s2 = -2.0 * c / e; ep.s3 = -1.0; ep.z = tkm2;
            }
            const bool last = (k == d);
            if (last && (d % 2 == 0)) {  // even degree: T_d > 0 at the wanted end -> flip so that it becomes the lowest
                ep.s1 = -ep.s1; ep.s2 = -ep.s2; ep.s3 = -ep.s3;
            }
            double* dst = last ? y : ctx->d_cheb[k % 3];
            const double* src = (k == 1) ? x : tkm1;
            ++hv_count;
            BH_TRY(bh_launch_hv(ctx, cJ, cU, cmu, kernel, src, dst, ep));
            if (k >= 2) tkm2 = tkm1;
            tkm1 = dst;
        }
        return (int)BH_OK;
    };
    BhSolve s2;
    ctx->cheb_stall_iter = quick ? 2 : 0;
    rc = lanczos_core(ctx, cheb, true, nev, ncv, tol, quick ? std::min(maxit, 80) : maxit, &s2);
    ctx->cheb_stall_iter = 0;
    const int restarts = s1.info.nrestart + s2.info.nrestart;
    if (rc == BH_ERR_NOCONV && quick) {  // a misplaced cut stalls the filtered iteration: full stage 1
        *retry = true;
        return BH_OK;
    }
    if (rc == BH_ERR_NOCONV) {  // the filtered iteration stalled: fall back to the reference algorithm
        rc = lanczos_core(ctx, plain, false, nev, ncv, tol, maxit, out);
        out->info.nmatvec = hv_count;
        out->info.nrestart += restarts;
        return rc;
    }
    if (rc != BH_OK) {
        *out = s2;
        out->info.nmatvec = hv_count;
        out->info.nrestart = restarts;
        return rc;
    }
    // ---- stage 3: Rayleigh-Ritz of H on span(V) ----
    {
        const int64_t n = ctx->nloc, ld = ctx->ld;
        const int G = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks(n, VEC_THREADS), (int64_t)ctx->sm_count * 4));
        const bool dist = ctx->partitioned && ctx->world > 1;
        const bool gram = !dist && ncv <= GRAM_N && ctx->rr_gram;
        if (gram) {
            // W = H V column by column, then M = V^T W in one pass over both (k_gram)
            const int nparts = ctx->sm_count * 2;
            if (ctx->hv_block_cols < ncv) {
                if (ctx->d_hv_block) cudaFree(ctx->d_hv_block);
                ctx->d_hv_block = nullptr;
                BH_CUDA(ctx, cudaMalloc(&ctx->d_hv_block, sizeof(double) * (size_t)ld * ncv));
                ctx->hv_block_cols = ncv;
            }
            if (!ctx->d_gram_part) BH_CUDA(ctx, cudaMalloc(&ctx->d_gram_part, sizeof(double) * (size_t)nparts * GRAM_N * GRAM_N));
            for (int col = 0; col < ncv; ++col) BH_TRY(plain(ctx->d_V + (int64_t)col * ld, ctx->d_hv_block + (int64_t)col * ld));
            BhProfScope prof(ctx, BH_PROF_GRAM, 16.0 * (double)n * ncv);
            k_gram<<<nparts, 256, 0, ctx->stream>>>(n, ld, ctx->d_V, ctx->d_hv_block, ncv, ctx->d_gram_part);
            k_gram_reduce<<<(ncv * ncv + 255) / 256, 256, 0, ctx->stream>>>(nparts, ctx->d_gram_part, ncv, ctx->d_small);
            ctx->launches += 2;
        }
        for (int col = 0; col < ncv && !gram; ++col) {
            BH_TRY(plain(ctx->d_V + (int64_t)col * ld, ctx->d_w));
            k_gemv_t<true><<<G, VEC_THREADS, 0, ctx->stream>>>(n, ld, ctx->d_V, 0, ncv, ctx->d_w, ctx->d_scal, 0, 0, ctx->d_part,
                                                              ctx->d_counter, 1);
            BH_LAUNCHED(ctx);
            if (dist) BH_TRY(bh_dist_allreduce_sum(ctx, ctx->d_scal + S_VF, ncv));
            BH_CUDA(ctx, cudaMemcpyAsync(ctx->d_small + (size_t)col * ncv, ctx->d_scal + S_VF, sizeof(double) * ncv,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
        }
        std::vector<double> M((size_t)ncv * ncv), evH, Z;
        BH_D2H(ctx, M.data(), ctx->d_small, sizeof(double) * (size_t)ncv * ncv);
        BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int a = 0; a < ncv; ++a)
            for (int b = a + 1; b < ncv; ++b) {
                const double v = 0.5 * (M[a + (size_t)b * ncv] + M[b + (size_t)a * ncv]);
                M[a + (size_t)b * ncv] = M[b + (size_t)a * ncv] = v;
            }
        bh_sym_eig(ncv, M, evH, Z);
        if (quick && !(evH[nev - 1] < cut - 0.02 * (cut - evH[0]))) {  // the cut was not safely above the nev-th level
            *retry = true;
            return BH_OK;
        }
        if (!(evH[nev - 1] < cut)) {
            // the wanted levels were not all below the cut (cannot happen by interlacing; kept as a safety net)
            rc = lanczos_core(ctx, plain, false, nev, ncv, tol, maxit, out);
            out->info.nmatvec = hv_count;
            return rc;
        }
        out->nev = nev;
        out->ncv = ncv;
        out->evals.assign(evH.begin(), evH.begin() + nev);
        out->Y.assign(Z.begin(), Z.begin() + (size_t)ncv * nev);
    }
    out->info = s2.info;
    out->info.nconv = nev;
    out->info.nmatvec = hv_count;
    out->info.nrestart = restarts;
    out->info.nreorth = s1.info.nreorth + s2.info.nreorth;
    return BH_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Solver driver.  Plain mode (cheb_degree = 1) is the reference's algorithm step for step.  The accelerated
// mode runs the same thick-restart Lanczos on a Chebyshev polynomial of H:
//   stage 1  a few plain restart cycles give Ritz values theta_0 <= ... (upper bounds of the eigenvalues by
//            interlacing), so cut = theta_{nev-1} + margin is guaranteed to lie above the nev wanted levels;
//   stage 2  Lanczos on B = -/+T_d((H - c)/e), [cut, hi] -> [-1, 1] with hi the Gershgorin bound: the wanted end
//            is amplified like cosh(d acosh(.)) and every Lanczos step (one re-orthogonalisation against the
//            basis, the dominant cost) advances the Krylov polynomial by d degrees.  The d applications of H use
//            the fused epilogue of the H.v kernels (no extra vector passes);
//   stage 3  Rayleigh-Ritz of H itself on the final basis (ncv H.v + ncv multi-dots, host 41 x 41 eigenproblem):
//            eigenvalues and vectors of H, not of the polynomial.
// Measured (numpy model, m=n=8): d = 7 needs 4.5x fewer Lanczos steps for 1.4-1.6x more H.v; eigenvalues agree
// with the plain solver to 1e-12.
// ---------------------------------------------------------------------------------------------------------
int bh_lanczos(bh_ctx* ctx, double cJ, double cU, double cmu, int nev, int ncv, double tol, int maxit, int kernel,
               BhSolve* out)
{
    const auto t_start = std::chrono::steady_clock::now();
    int hv_count = 0;
    BH_TRY(bh_ensure_workspace(ctx, std::min(ncv, BH_MAX_NCV)));  // full size up front: no reallocation between the stages
    LanczosOp plain = [&](const double* x, double* y) {
        ++hv_count;
        if (ctx->parent && ctx->parent->hub && ctx->parent->batch_plain) {
            // lockstep solve: a plain H.v is a filter request of degree 1 with c = 0, e = 1 (s1 = 1, s2 = -0 is skipped)
            bool handled = false;
            BH_TRY(bh_batch_filter(ctx, x, y, 0.0, 1.0, cJ, cU, cmu, 1, &handled));
            if (handled) return (int)BH_OK;
        }
        return bh_launch_hv(ctx, cJ, cU, cmu, kernel, x, y);
    };
    const int d = ctx->cheb_degree;
    static const int64_t accel_min_d = getenv("BH_ACCEL_MIN_D") ? atoll(getenv("BH_ACCEL_MIN_D")) : 100;  // below: plain mode (r02: raised coverage from D >= 2000, see DESIGN.md "Multiplicities")
    const bool accel = d > 1 && !ctx->user_matrix && kernel != BH_HV_USER && ctx->D >= accel_min_d && ncv >= nev + 2 && ncv <= ctx->D;
    if (!accel) {
        const int rc = lanczos_core(ctx, plain, false, nev, ncv, tol, maxit, out);
        out->info.nmatvec = hv_count;
        return rc;
    }
    // Stage 1 has two forms.  Quick (large systems): ONE cycle of `cheb_quick` plain steps gives theta_0 >= E_0 and the
    // cut is placed by the spectrum-fraction rule alone; stage 2 starts from that cycle's ground Ritz vector.  The wanted
    // levels must lie below the cut: that is verified on the final Rayleigh-Ritz values, and when it does not hold (the nev
    // levels span more than cheb_frac of the spectrum: small systems) the solve is repeated with the full form -- cheb_pre
    // restart cycles of the reference algorithm whose nev-th Ritz value bounds the nev-th level from above (interlacing).
    // Model (tools/model_block_lanczos.py, m = n = 10): the quick form needs 2-15 % fewer H.v in total and saves ~80 H.v
    // and ~50 full re-orthogonalisation steps of stage 1 per grid point.
    const bool quick_ok = ctx->cheb_quick >= 4 && ctx->D >= 50000 && ncv > ctx->cheb_quick + 2;
    for (int attempt = quick_ok ? 0 : 1; attempt < 2; ++attempt) {
        const bool quick = attempt == 0;
        bool retry = false;
        const int rc = accel_solve(ctx, plain, hv_count, cJ, cU, cmu, nev, ncv, tol, maxit, kernel, quick, &retry, out);
        if (!retry) {
            out->info.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
            return rc;
        }
    }
    return bh_fail(ctx, BH_ERR_STATE, "bh_lanczos: unreachable");
}

int bh_ritz_vector(bh_ctx* ctx, const BhSolve& s, int col, double* x_dev)
{
    BH_H2D(ctx, ctx->d_small, s.Y.data() + (size_t)col * s.ncv, sizeof(double) * s.ncv);
    const int64_t n = ctx->nloc;
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks(n, VEC_THREADS), (int64_t)ctx->sm_count * 4));
    k_lincomb<<<G, VEC_THREADS, 0, ctx->stream>>>(n, ctx->ld, ctx->d_V, s.ncv, ctx->d_small, x_dev);
    if (ctx->partitioned && ctx->world > 1) {
        k_norm<<<G, VEC_THREADS, 0, ctx->stream>>>(n, x_dev, ctx->d_scal, S_FLAG + 3, ctx->d_part, ctx->d_counter, 1);
        BH_TRY(bh_dist_allreduce_sum(ctx, ctx->d_scal + S_FLAG + 3, 1));
        k_fin_sqrt<<<1, 1, 0, ctx->stream>>>(ctx->d_scal, S_FLAG + 3, S_VF);
    } else {
        k_norm<<<G, VEC_THREADS, 0, ctx->stream>>>(n, x_dev, ctx->d_scal, S_VF, ctx->d_part, ctx->d_counter);
    }
    k_scale_inplace<<<G, VEC_THREADS, 0, ctx->stream>>>(n, x_dev, ctx->d_scal, S_VF);
    ctx->launches += 3;
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}

extern "C" int bh_eigs(bh_ctx* ctx, double cJ, double cU, double cmu, int nev, int ncv, double tol, int maxit, int kernel,
                       int order, double* evals, double* evecs, bh_eigs_info* info)
{
    if (!ctx || !ctx->D) return bh_fail(ctx, BH_ERR_STATE, "bh_eigs: call bh_setup first");
    if (!evals || order < 0 || order > 2) return bh_fail(ctx, BH_ERR_ARG, "bh_eigs: bad argument");
    if (ctx->partitioned && order != BH_ORDER_LEX) return bh_fail(ctx, BH_ERR_ARG, "bh_eigs: a row-partitioned context returns LEX-order slices only");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    BhSolve s;
    const int rc = bh_lanczos(ctx, cJ, cU, cmu, nev, ncv, tol, maxit, kernel, &s);
    if (info) *info = s.info;
    if (rc != BH_OK && rc != BH_ERR_NOCONV) return rc;
    for (size_t i = 0; i < s.evals.size(); ++i) evals[i] = s.evals[i];
    if (evecs && !s.evals.empty()) {
        BH_TRY(bh_ensure_staging(ctx));
        for (int c = 0; c < nev; ++c) {
            BH_TRY(bh_ritz_vector(ctx, s, c, ctx->d_x));
            BH_TRY(bh_permute_vec(ctx, order, true, ctx->d_x, ctx->d_y));
            BH_D2H(ctx, evecs + (size_t)c * ctx->nloc, ctx->d_y, sizeof(double) * ctx->nloc);
            BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return rc;
}
