// small.cu -- many grid points of a SMALL system at once: one CTA (one SM) per grid point.
//
// The reference sweeps the grid with an OpenMP team, one eigensolve per thread (src/analysis.cpp:302-343).  For the small
// configurations of BASELINE.json (C1: m = n = 8, D = 6 435; C2: m = n = 10, D = 92 378) one eigensolve cannot fill a B200:
// a Lanczos step is a few microseconds of work behind ~25 us of launch and grid-barrier latency (12.5 ms per
// m = n = 8 point).  Here the grid points map to CTAs instead: bh_points hands the whole list to bh_points_small, every
// point gets its own Krylov workspace in HBM, and ONE launch of k_small_cycle runs a complete restart cycle of every
// unfinished point -- thick restart V <- V Y, then for each new basis column the operator (H, or the degree-d Chebyshev
// filter of H, matrix-free chain rows), the three-term update and the full block Gram-Schmidt re-orthogonalisation -- with
// __syncthreads as the only barrier.  Between cycles the host reads the recurrence scalars of all points in one copy,
// solves the <= 62 x 62 projected problems on a few host threads (Spectra's convergence test and restart size,
// HermEigsBase.h:152-196, exactly as lanczos.cu) and uploads the Ritz coefficients.  The algorithm per point is the one
// of bh_lanczos: stage 1 plain cycles (or the quick form), stage 2 on the Chebyshev filter, stage 3 Rayleigh-Ritz of H;
// then ground-state Ritz vector, SPDM and the three output columns.  Points that break down (J = 0: invariant subspace)
// or fail a safety check are handed to the ordinary single-point path afterwards.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "device_utils.cuh"

#define SM_THREADS 1024
#define SM_NC 128  // stride of the scalar arrays (>= ncv + 2)
#define SM_ALPHA 0
#define SM_BETA SM_NC
#define SM_OFFD (2 * SM_NC)
#define SM_FLAG (3 * SM_NC)
#define SM_SCAL (3 * SM_NC + 8)
#define SM_CH 8       // basis columns per re-orthogonalisation block
#define SM_CROWS 128  // rows per shared-memory chunk of the restart product
#define SM_MAX_NCV 64

struct SmallDesc {
    double cJ, cU, cmu;  // coefficients of the three operators at this grid point
    double c, e;         // filter centre / half width (mode 1)
    int mode;            // 0: operator = H, 1: operator = -/+ T_d((H - c) / e)
    int d;
    int ncv;             // basis columns of this solve
    int from;            // first step of the cycle
    int compress_k;      // > 0: V[:, 0..k) <- V[:, 0..ncv) Y first; Y (ncv x k, column-major) in the point's Y buffer
    int init;            // 1: the start vector is in w: f = Op w, beta_0 = |f| before the first step
    int active;
    int pad;
};

struct SmallPtrs {
    double* V;     // point p: V + p * slot_V, (ncv_max + 1) columns of ld doubles
    double* vec;   // point p: vec + p * 5 ld: w, f, t0, t1, t2
    double* scal;  // point p: scal + p * SM_SCAL
    double* Y;     // point p: Y + p * SM_MAX_NCV^2
    int64_t ld, slot_V;
};

// sum of NV per-thread values over the CTA, fixed order; results in out[0..NV) (shared), valid after the call
template <int NV>
__device__ __forceinline__ void small_reduce(double (&v)[NV], double (*red)[SM_CH], double* out)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = bh_warp_sum(v[j]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) red[wid][j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int w = 0; w < SM_THREADS / 32; ++w) t += red[w][threadIdx.x];
        out[threadIdx.x] = t;
    }
    __syncthreads();
}

// y = s1 (H x) + s2 x + s3 z over all rows of one point (the fused epilogue of the H.v kernels in hv.cu, same arithmetic)
template <int M>
__device__ __forceinline__ void small_hv(const BhTables& t, int64_t D, const uint64_t* __restrict__ states, const double* __restrict__ dU,
                                         double cJ, double cU, double shift, const double* x, double* y, double s1, double s2, double s3,
                                         const double* z)
{
    for (int64_t r = threadIdx.x; r < D; r += SM_THREADS) {
        const uint64_t s = __ldg(states + r);
        const double acc = bh_chain_row_sum<M, true>(t, s, (int)r, x);
        const double diag = __dadd_rn(__dmul_rn(__ldg(dU + r), cU), shift);
        const double xv = x[r];
        double out = s1 * (diag * xv - (2.0 * cJ) * acc);
        if (s2 != 0.0) out = fma(s2, xv, out);
        if (z) out = fma(s3, z[r], out);
        y[r] = out;
    }
    __syncthreads();
}

// y = Op x  (mode 0: H; mode 1: the Chebyshev recurrence of bh_lanczos stage 2 through the buffers tb[0..2])
template <int M>
__device__ __forceinline__ void small_op(const BhTables& t, int64_t D, const uint64_t* __restrict__ states, const double* __restrict__ dU,
                                         const SmallDesc& de, double shift, const double* x, double* y, double* const* tb)
{
    if (de.mode == 0) {
        small_hv<M>(t, D, states, dU, de.cJ, de.cU, shift, x, y, 1.0, 0.0, 0.0, nullptr);
        return;
    }
    const double* tkm2 = x;
    const double* tkm1 = nullptr;
    for (int k = 1; k <= de.d; ++k) {
        double s1, s2, s3 = 0.0;
        const double* z = nullptr;
        if (k == 1) {
            s1 = 1.0 / de.e; s2 = -de.c / de.e;
        } else {
            s1 = 2.0 / de.e; s2 = -2.0 * de.c / de.e; s3 = -1.0; z = tkm2;
        }
        const bool last = (k == de.d);
        if (last && (de.d % 2 == 0)) { s1 = -s1; s2 = -s2; s3 = -s3; }
        double* dst = last ? y : tb[k % 3];
        const double* src = (k == 1) ? x : tkm1;
        small_hv<M>(t, D, states, dU, de.cJ, de.cU, shift, src, dst, s1, s2, s3, z);
        if (k >= 2) tkm2 = tkm1;
        tkm1 = dst;
    }
}

template <int M>
__global__ void __launch_bounds__(SM_THREADS, 1)
k_small_cycle(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states, const double* __restrict__ dU,
              const SmallDesc* __restrict__ descs, SmallPtrs P, double near0)
{
    extern __shared__ double dyn[];  // restart product: vt[ncv][SM_CROWS], ys[ncv][k]
    __shared__ BhTables t;
    __shared__ double red[SM_THREADS / 32][SM_CH];
    __shared__ double cs[SM_CH];
    const SmallDesc de = descs[blockIdx.x];
    if (!de.active) return;
    bh_stage_tables(&t, gtab);
    const int tid = threadIdx.x;
    const int64_t ld = P.ld;
    double* V = P.V + (int64_t)blockIdx.x * P.slot_V;
    double* w = P.vec + (int64_t)blockIdx.x * 5 * ld;
    double* f = w + ld;
    double* tb[3] = {w + 2 * ld, w + 3 * ld, w + 4 * ld};
    double* scal = P.scal + (int64_t)blockIdx.x * SM_SCAL;
    const double* Yp = P.Y + (int64_t)blockIdx.x * SM_MAX_NCV * SM_MAX_NCV;
    const int ncv = de.ncv;
    const double shift = __dmul_rn(-(double)t.n, de.cmu);

    // ---- thick restart: V[:, 0..k) <- V Y, in place, SM_CROWS rows at a time through shared memory ----
    if (de.compress_k > 0) {
        const int k = de.compress_k;
        double* vt = dyn;
        double* ys = dyn + (size_t)ncv * SM_CROWS;
        for (int idx = tid; idx < ncv * k; idx += SM_THREADS) {
            const int i = idx / k, j = idx % k;
            ys[idx] = Yp[i + (size_t)j * ncv];
        }
        const int rr = tid % SM_CROWS, cg = tid / SM_CROWS;
        for (int64_t r0 = 0; r0 < D; r0 += SM_CROWS) {
            __syncthreads();
            for (int idx = tid; idx < ncv * SM_CROWS; idx += SM_THREADS) {
                const int i = idx / SM_CROWS, r = idx % SM_CROWS;
                vt[idx] = (r0 + r < D) ? V[(int64_t)i * ld + r0 + r] : 0.0;
            }
            __syncthreads();
            if (r0 + rr < D)
                for (int j = cg; j < k; j += SM_THREADS / SM_CROWS) {
                    double acc = 0.0;
                    for (int i = 0; i < ncv; ++i) acc = fma(vt[i * SM_CROWS + rr], ys[i * k + j], acc);
                    V[(int64_t)j * ld + r0 + rr] = acc;
                }
        }
        __syncthreads();
        if (tid == 0) scal[SM_BETA + k] = scal[SM_BETA + ncv];  // the residual f and its norm carry over
        __syncthreads();
    }

    // ---- start: f = Op w, beta_0 = |f| ----
    if (de.init) {
        small_op<M>(t, D, states, dU, de, shift, w, f, tb);
        double a[1] = {0.0};
        for (int64_t r = tid; r < D; r += SM_THREADS) a[0] = fma(f[r], f[r], a[0]);
        small_reduce<1>(a, red, cs);
        if (tid == 0) scal[SM_BETA + 0] = sqrt(cs[0]);
        __syncthreads();
    }

    for (int i = de.from; i < ncv; ++i) {
        const double beta = scal[SM_BETA + i];
        if (!(beta > near0)) {  // invariant subspace: the host hands this point to the single-point path
            if (tid == 0) {
                scal[SM_FLAG] = 1.0;
                scal[SM_FLAG + 1] = (double)i;
            }
            break;
        }
        double* vi = V + (int64_t)i * ld;
        const double inv = 1.0 / beta;
        for (int64_t r = tid; r < D; r += SM_THREADS) vi[r] = f[r] * inv;
        __syncthreads();
        small_op<M>(t, D, states, dU, de, shift, vi, w, tb);
        const bool far = de.from > 0 && i == de.from;  // first step after a restart: couples to every kept Ritz vector
        const bool sub = i > 0 && !far;
        const double* vp = V + (int64_t)(i > 0 ? i - 1 : 0) * ld;
        // three-term update and alpha (every thread owns the same rows in every vector operation below)
        {
            double a[1] = {0.0};
            for (int64_t r = tid; r < D; r += SM_THREADS) {
                double wr = w[r];
                if (sub) wr = fma(-beta, vp[r], wr);
                f[r] = wr;
                a[0] = fma(vi[r], wr, a[0]);
            }
            small_reduce<1>(a, red, cs);
        }
        double alpha = cs[0], offd = sub ? beta : 0.0;
        for (int64_t r = tid; r < D; r += SM_THREADS) f[r] = fma(-alpha, vi[r], f[r]);
        // block modified Gram-Schmidt against columns 0..i (Lanczos.h:150-171: the corrections of columns i and i-1 go into T)
        const int passes = far ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass)
            for (int c0 = 0; c0 <= i; c0 += SM_CH) {
                const int nc = min(SM_CH, i + 1 - c0);
                const double* VB = V + (int64_t)c0 * ld;
                double acc[SM_CH];
#pragma unroll
                for (int j = 0; j < SM_CH; ++j) acc[j] = 0.0;
                for (int64_t r = tid; r < D; r += SM_THREADS) {
                    const double fr = f[r];
#pragma unroll
                    for (int j = 0; j < SM_CH; ++j)
                        if (j < nc) acc[j] = fma(VB[(int64_t)j * ld + r], fr, acc[j]);
                }
                small_reduce<SM_CH>(acc, red, cs);
                double c[SM_CH];
#pragma unroll
                for (int j = 0; j < SM_CH; ++j) c[j] = (j < nc) ? cs[j] : 0.0;
                if (i >= c0 && i < c0 + SM_CH) alpha += cs[i - c0];
                if (sub && i - 1 >= c0 && i - 1 < c0 + SM_CH) offd += cs[i - 1 - c0];
                for (int64_t r = tid; r < D; r += SM_THREADS) {
                    double fr = f[r];
#pragma unroll
                    for (int j = 0; j < SM_CH; ++j)
                        if (j < nc) fr = fma(-VB[(int64_t)j * ld + r], c[j], fr);
                    f[r] = fr;
                }
                __syncthreads();  // cs is rewritten by the next block
            }
        {
            double a[1] = {0.0};
            for (int64_t r = tid; r < D; r += SM_THREADS) a[0] = fma(f[r], f[r], a[0]);
            small_reduce<1>(a, red, cs);
        }
        if (tid == 0) {
            scal[SM_ALPHA + i] = alpha;
            scal[SM_OFFD + i] = offd;
            scal[SM_BETA + i + 1] = sqrt(cs[0]);
        }
        __syncthreads();
    }
}

// Rayleigh-Ritz of H on span(V): M[:, j] = V^T (H v_j), written to the point's Y buffer (ncv x ncv, column-major)
template <int M>
__global__ void __launch_bounds__(SM_THREADS, 1)
k_small_rr(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states, const double* __restrict__ dU,
           const SmallDesc* __restrict__ descs, SmallPtrs P)
{
    __shared__ BhTables t;
    __shared__ double red[SM_THREADS / 32][SM_CH];
    __shared__ double cs[SM_CH];
    const SmallDesc de = descs[blockIdx.x];
    if (!de.active) return;
    bh_stage_tables(&t, gtab);
    const int tid = threadIdx.x;
    const int64_t ld = P.ld;
    double* V = P.V + (int64_t)blockIdx.x * P.slot_V;
    double* w = P.vec + (int64_t)blockIdx.x * 5 * ld;
    double* Mo = P.Y + (int64_t)blockIdx.x * SM_MAX_NCV * SM_MAX_NCV;
    const int ncv = de.ncv;
    const double shift = __dmul_rn(-(double)t.n, de.cmu);
    for (int j = 0; j < ncv; ++j) {
        small_hv<M>(t, D, states, dU, de.cJ, de.cU, shift, V + (int64_t)j * ld, w, 1.0, 0.0, 0.0, nullptr);
        for (int c0 = 0; c0 < ncv; c0 += SM_CH) {
            const int nc = min(SM_CH, ncv - c0);
            const double* VB = V + (int64_t)c0 * ld;
            double acc[SM_CH];
#pragma unroll
            for (int q = 0; q < SM_CH; ++q) acc[q] = 0.0;
            for (int64_t r = tid; r < D; r += SM_THREADS) {
                const double wr = w[r];
#pragma unroll
                for (int q = 0; q < SM_CH; ++q)
                    if (q < nc) acc[q] = fma(VB[(int64_t)q * ld + r], wr, acc[q]);
            }
            small_reduce<SM_CH>(acc, red, cs);
            if (tid < nc) Mo[(c0 + tid) + (size_t)j * ncv] = cs[tid];
            __syncthreads();
        }
    }
}

// w = V[:, 0..cnt) y (y = the first cnt doubles of the point's Y buffer); normalise = 1: w /= |w|  (Ritz vector)
__global__ void __launch_bounds__(SM_THREADS, 1)
k_small_lincomb(int64_t D, const SmallDesc* __restrict__ descs, SmallPtrs P, int normalise)
{
    __shared__ double coef[SM_MAX_NCV];
    __shared__ double red[SM_THREADS / 32][SM_CH];
    __shared__ double cs[SM_CH];
    const SmallDesc de = descs[blockIdx.x];
    if (!de.active) return;
    const int tid = threadIdx.x;
    const int64_t ld = P.ld;
    const double* V = P.V + (int64_t)blockIdx.x * P.slot_V;
    double* w = P.vec + (int64_t)blockIdx.x * 5 * ld;
    const double* y = P.Y + (int64_t)blockIdx.x * SM_MAX_NCV * SM_MAX_NCV;
    const int cnt = de.ncv;
    if (tid < cnt) coef[tid] = y[tid];
    __syncthreads();
    double a[1] = {0.0};
    for (int64_t r = tid; r < D; r += SM_THREADS) {
        double acc = 0.0;
        for (int j = 0; j < cnt; ++j) acc += V[(int64_t)j * ld + r] * coef[j];
        w[r] = acc;
        a[0] = fma(acc, acc, a[0]);
    }
    if (!normalise) return;
    small_reduce<1>(a, red, cs);
    const double inv = 1.0 / sqrt(cs[0]);
    for (int64_t r = tid; r < D; r += SM_THREADS) w[r] *= inv;
}

// Gershgorin bounds of every point: out[2 p] = upper bound, out[2 p + 1] = lower bound
__global__ void __launch_bounds__(SM_THREADS, 1)
k_small_gersh(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states, const double* __restrict__ dU,
              const SmallDesc* __restrict__ descs, double* __restrict__ out)
{
    __shared__ BhTables t;
    __shared__ double shi[SM_THREADS / 32], slo[SM_THREADS / 32];
    const SmallDesc de = descs[blockIdx.x];
    bh_stage_tables(&t, gtab);
    double hi = -1e300, lo = 1e300;
    for (int64_t r = threadIdx.x; r < D; r += SM_THREADS) {
        const uint64_t s = states[r];
        double off = 0.0;
        for (int b = 0; b < t.nbonds; ++b) {
            const int bd = t.bond[b];
            const int dst = bd & 15, src = (bd >> 4) & 15, wgt = bd >> 8;
            off += (double)wgt * t.sq[(bh_occ(s, dst) + 1) * bh_occ(s, src)];
        }
        off *= fabs(de.cJ);
        const double diag = dU[r] * de.cU - (double)t.n * de.cmu;
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
        for (int q = 1; q < SM_THREADS / 32; ++q) { hi = fmax(hi, shi[q]); lo = fmin(lo, slo[q]); }
        out[2 * blockIdx.x] = hi;
        out[2 * blockIdx.x + 1] = lo;
    }
}

typedef void (*small_cycle_fn)(const BhTables*, int64_t, const uint64_t*, const double*, const SmallDesc*, SmallPtrs, double);
typedef void (*small_rr_fn)(const BhTables*, int64_t, const uint64_t*, const double*, const SmallDesc*, SmallPtrs);
static small_cycle_fn small_cycle_kernel(int m)
{
    switch (m) {
        case 3: return k_small_cycle<3>;
        case 4: return k_small_cycle<4>;
        case 5: return k_small_cycle<5>;
        case 6: return k_small_cycle<6>;
        case 7: return k_small_cycle<7>;
        case 8: return k_small_cycle<8>;
        case 9: return k_small_cycle<9>;
        case 10: return k_small_cycle<10>;
        case 11: return k_small_cycle<11>;
        case 12: return k_small_cycle<12>;
        case 13: return k_small_cycle<13>;
        case 14: return k_small_cycle<14>;
        case 15: return k_small_cycle<15>;
        case 16: return k_small_cycle<16>;
    }
    return nullptr;
}
static small_rr_fn small_rr_kernel(int m)
{
    switch (m) {
        case 3: return k_small_rr<3>;
        case 4: return k_small_rr<4>;
        case 5: return k_small_rr<5>;
        case 6: return k_small_rr<6>;
        case 7: return k_small_rr<7>;
        case 8: return k_small_rr<8>;
        case 9: return k_small_rr<9>;
        case 10: return k_small_rr<10>;
        case 11: return k_small_rr<11>;
        case 12: return k_small_rr<12>;
        case 13: return k_small_rr<13>;
        case 14: return k_small_rr<14>;
        case 15: return k_small_rr<15>;
        case 16: return k_small_rr<16>;
    }
    return nullptr;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct bh_small_ws {
    int64_t cap_points = 0, ld = 0;
    int ncv_max = 0;
    double *V = nullptr, *vec = nullptr, *scal = nullptr, *Y = nullptr, *bounds = nullptr, *start = nullptr, *rho = nullptr, *rho_part = nullptr;
    SmallDesc* d_desc = nullptr;
};

void bh_small_release(bh_ctx* ctx)
{
    bh_small_ws* ws = static_cast<bh_small_ws*>(ctx->small_ws);
    if (!ws) return;
    cudaFree(ws->V); cudaFree(ws->vec); cudaFree(ws->scal); cudaFree(ws->Y); cudaFree(ws->bounds); cudaFree(ws->start);
    cudaFree(ws->rho); cudaFree(ws->rho_part); cudaFree(ws->d_desc);
    delete ws;
    ctx->small_ws = nullptr;
}

bool bh_small_supported(const bh_ctx* ctx, int kernel, int64_t npoints, int nb_eigen)
{
    static const int enabled = getenv("BH_SMALL") ? atoi(getenv("BH_SMALL")) : 1;
    static const int min_points = getenv("BH_SMALL_MIN_POINTS") ? atoi(getenv("BH_SMALL_MIN_POINTS")) : 8;
    return enabled && kernel == BH_HV_MATRIX_FREE && !ctx->user_matrix && !ctx->partitioned && !ctx->parent && ctx->h_tab.chain == 2 &&
           ctx->m >= 3 && ctx->D <= 100000 && npoints >= min_points && 2 * nb_eigen + 1 <= SM_MAX_NCV - 2 && 2 * nb_eigen + 1 <= ctx->D &&
           nb_eigen <= ctx->D - 2;
}

namespace {

struct SmallPoint {
    int stage = 1;  // 1 plain cycles, 2 filtered cycles, 3 Rayleigh-Ritz pending, 4 Ritz vector pending, 5 done, 6 single-point path
    int nev = 0, ncv = 0, maxit = 0;  // of the current stage
    int from = 0, iter = 0, init = 1, compress_k = 0;
    bool quick = false;
    std::vector<double> theta, coup, evals, Y;
    std::vector<double> final_evals, final_y;
    double hi = 0, lo = 0, cut = 0, c = 0, e = 0;
    int nmatvec = 0, nrestart = 0, nsteps = 0;
};

template <class F>
void parallel_for(int n, F&& body)
{
    const int nt = std::max(1, std::min<int>({n, 16, (int)std::thread::hardware_concurrency()}));
    if (nt == 1) {
        for (int i = 0; i < n; ++i) body(i);
        return;
    }
    std::vector<std::thread> th;
    std::atomic<int> next(0);
    for (int q = 0; q < nt; ++q)
        th.emplace_back([&] {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i);
        });
    for (auto& x : th) x.join();
}

}  // namespace

// defined in observables.cu: rho of npts vectors (phi_p = phi + p * stride) in one launch pair
int bh_spdm_batch_dev(bh_ctx* ctx, const double* phi, int64_t stride, int npts, int ncols, double* d_part, double* d_rho, double* rho_host);

int bh_points_small(bh_ctx* ctx, int64_t npoints, const double* cJ, const double* cU, const double* cmu, int nb_eigen, double* out3,
                    bh_eigs_info* infos)
{
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const auto t_start = std::chrono::steady_clock::now();
    const int m = ctx->m, nev = nb_eigen, ncv = 2 * nb_eigen + 1;
    const int64_t D = ctx->D, ld = ctx->ld;
    const double tol = 1e-10;
    const int maxit = 1000;
    const double eps = std::numeric_limits<double>::epsilon();
    const double near0 = std::numeric_limits<double>::min() * 10.0;
    const double eps23 = std::pow(eps, 2.0 / 3.0);
    const int d = ctx->cheb_degree;
    static const int64_t accel_min_d = getenv("BH_ACCEL_MIN_D") ? atoll(getenv("BH_ACCEL_MIN_D")) : 100;
    const bool accel = d > 1 && D >= accel_min_d;
    const bool quick = accel && ctx->cheb_quick >= 4 && D >= 50000 && ncv > ctx->cheb_quick + 2;
    cudaStream_t st = ctx->stream;
    small_cycle_fn cycle = small_cycle_kernel(m);
    small_rr_fn rrk = small_rr_kernel(m);
    if (!cycle || !rrk) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "small-system solver: unsupported chain length");

    // chunks bounded by memory (8 GB of Krylov workspace)
    const int64_t per_point = sizeof(double) * (size_t)ld * (ncv + 1 + 5);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>({npoints, (int64_t)(((size_t)8 << 30) / per_point), (int64_t)4096}));

    bh_small_ws* ws = static_cast<bh_small_ws*>(ctx->small_ws);
    if (!ws || ws->cap_points < chunk || ws->ld != ld || ws->ncv_max < ncv) {
        bh_small_release(ctx);
        ws = new bh_small_ws();
        ctx->small_ws = ws;
        ws->cap_points = chunk;
        ws->ld = ld;
        ws->ncv_max = ncv;
        BH_CUDA(ctx, cudaMalloc(&ws->V, sizeof(double) * (size_t)chunk * ld * (ncv + 1)));
        BH_CUDA(ctx, cudaMalloc(&ws->vec, sizeof(double) * (size_t)chunk * ld * 5));
        BH_CUDA(ctx, cudaMalloc(&ws->scal, sizeof(double) * (size_t)chunk * SM_SCAL));
        BH_CUDA(ctx, cudaMalloc(&ws->Y, sizeof(double) * (size_t)chunk * SM_MAX_NCV * SM_MAX_NCV));
        BH_CUDA(ctx, cudaMalloc(&ws->bounds, sizeof(double) * (size_t)chunk * 2));
        BH_CUDA(ctx, cudaMalloc(&ws->start, sizeof(double) * (size_t)ld));
        BH_CUDA(ctx, cudaMalloc(&ws->rho, sizeof(double) * (size_t)chunk * m * m));
        BH_CUDA(ctx, cudaMalloc(&ws->rho_part, sizeof(double) * (size_t)chunk * 8 * m * m));
        BH_CUDA(ctx, cudaMalloc(&ws->d_desc, sizeof(SmallDesc) * (size_t)chunk));
        BH_CUDA(ctx, cudaMemsetAsync(ws->start, 0, sizeof(double) * (size_t)ld, st));
        BH_TRY(bh_lcg_fill_dev(ctx, ws->start, D));  // Spectra's start vector, the same for every point
    }
    SmallPtrs P{ws->V, ws->vec, ws->scal, ws->Y, ld, (int64_t)ld * (ncv + 1)};
    const size_t dyn_smem = sizeof(double) * ((size_t)ncv * SM_CROWS + (size_t)ncv * ncv);
    BH_CUDA(ctx, cudaFuncSetAttribute(cycle, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    if (dyn_smem > 160 * 1024) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "small-system solver: ncv too large");

    std::vector<int64_t> later;  // points for the single-point path
    const bool verbose = getenv("BH_BATCH_VERBOSE") != nullptr;
    double t_gpu_wait = 0, t_host = 0, t_setup = 0, t_final = 0;
    int ncycles_total = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs_since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    for (int64_t p0 = 0; p0 < npoints; p0 += chunk) {
        const int np = (int)std::min<int64_t>(chunk, npoints - p0);
        const auto t_chunk = now();
        std::vector<SmallPoint> pts(np);
        std::vector<SmallDesc> desc(np);
        std::vector<double> h_scal((size_t)np * SM_SCAL), h_Y((size_t)np * SM_MAX_NCV * SM_MAX_NCV), h_bounds((size_t)np * 2);
        BH_CUDA(ctx, cudaMemsetAsync(ws->scal, 0, sizeof(double) * (size_t)np * SM_SCAL, st));
        for (int p = 0; p < np; ++p) {
            SmallPoint& sp = pts[p];
            sp.quick = quick;
            if (!accel) { sp.nev = nev; sp.ncv = ncv; sp.maxit = maxit; }
            else if (quick) { sp.nev = 1; sp.ncv = ctx->cheb_quick; sp.maxit = 0; }
            else { sp.nev = nev; sp.ncv = ncv; sp.maxit = std::min(maxit, ctx->cheb_pre); }
            SmallDesc& de = desc[p];
            std::memset(&de, 0, sizeof(de));
            de.cJ = cJ[p0 + p]; de.cU = cU[p0 + p]; de.cmu = cmu[p0 + p];
            de.d = d;
            BH_CUDA(ctx, cudaMemcpyAsync(ws->vec + (size_t)p * 5 * ld, ws->start, sizeof(double) * (size_t)ld, cudaMemcpyDeviceToDevice, st));
        }
        BH_H2D(ctx, ws->d_desc, desc.data(), sizeof(SmallDesc) * np);
        if (accel) {
            k_small_gersh<<<np, SM_THREADS, 0, st>>>(ctx->d_tab, D, ctx->d_states, ctx->d_dU, ws->d_desc, ws->bounds);
            BH_LAUNCHED(ctx);
            BH_D2H(ctx, h_bounds.data(), ws->bounds, sizeof(double) * 2 * np);
            BH_CUDA(ctx, cudaStreamSynchronize(st));
            for (int p = 0; p < np; ++p) { pts[p].hi = h_bounds[2 * p]; pts[p].lo = h_bounds[2 * p + 1]; }
        }

        t_setup += secs_since(t_chunk);
        for (;;) {
            // ---- what does every point do next? ----
            int ncycle = 0, nrr = 0, nlin = 0;
            for (int p = 0; p < np; ++p) ncycle += (pts[p].stage == 1 || pts[p].stage == 2);
            if (ncycle) {
                for (int p = 0; p < np; ++p) {
                    SmallPoint& sp = pts[p];
                    SmallDesc& de = desc[p];
                    de.active = (sp.stage == 1 || sp.stage == 2);
                    if (!de.active) continue;
                    de.mode = (sp.stage == 2) ? 1 : 0;
                    de.c = sp.c; de.e = sp.e;
                    de.ncv = sp.ncv; de.from = sp.from; de.compress_k = sp.compress_k; de.init = sp.init;
                    if (sp.compress_k > 0)
                        std::copy(sp.Y.begin(), sp.Y.begin() + (size_t)sp.ncv * sp.compress_k, h_Y.begin() + (size_t)p * SM_MAX_NCV * SM_MAX_NCV);
                }
                BH_H2D(ctx, ws->d_desc, desc.data(), sizeof(SmallDesc) * np);
                BH_H2D(ctx, ws->Y, h_Y.data(), sizeof(double) * h_Y.size());
                {
                    BhProfScope prof(ctx, BH_PROF_SMALL, 0.0);
                    cycle<<<np, SM_THREADS, dyn_smem, st>>>(ctx->d_tab, D, ctx->d_states, ctx->d_dU, ws->d_desc, P, near0);
                }
                BH_LAUNCHED(ctx);
                BH_CUDA(ctx, cudaGetLastError());
                const auto t_w = now();
                BH_D2H(ctx, h_scal.data(), ws->scal, sizeof(double) * (size_t)np * SM_SCAL);
                BH_CUDA(ctx, cudaStreamSynchronize(st));
                t_gpu_wait += secs_since(t_w);
                ++ncycles_total;
                const auto t_h = now();
                // ---- Ritz pairs of every projected matrix, Spectra's convergence test and restart size (host threads) ----
                parallel_for(np, [&](int p) {
                    SmallPoint& sp = pts[p];
                    if (!(sp.stage == 1 || sp.stage == 2)) return;
                    const double* hs = h_scal.data() + (size_t)p * SM_SCAL;
                    const int nv = sp.ncv, ne = sp.nev;
                    const int opcost = (sp.stage == 2) ? d : 1;
                    if (hs[SM_FLAG] != 0.0) { sp.stage = 6; return; }
                    sp.nmatvec += opcost * ((sp.init ? 1 : 0) + (nv - sp.from));
                    sp.nsteps += nv - sp.from;
                    std::vector<double> T((size_t)nv * nv, 0.0), evals, Y;
                    for (int l = 0; l < sp.from; ++l) {
                        T[l + (size_t)l * nv] = sp.theta[l];
                        T[sp.from + (size_t)l * nv] = T[l + (size_t)sp.from * nv] = sp.coup[l];
                    }
                    for (int i = sp.from; i < nv; ++i) {
                        T[i + (size_t)i * nv] = hs[SM_ALPHA + i];
                        if (i > sp.from) T[i + (size_t)(i - 1) * nv] = T[(i - 1) + (size_t)i * nv] = hs[SM_OFFD + i];
                    }
                    bh_sym_eig(nv, T, evals, Y);
                    const double beta_last = hs[SM_BETA + nv];
                    int nconv = 0;
                    for (int l = 0; l < ne; ++l) {
                        const double thresh = tol * std::max(eps23, std::fabs(evals[l]));
                        nconv += (std::fabs(Y[(nv - 1) + (size_t)l * nv]) * beta_last < thresh);
                    }
                    sp.init = 0;
                    // misplaced cut of the quick form (lanczos_core has the same rule): the nev-th Ritz value of the filtered
                    // operator is still inside the damped band after three cycles -> the single-point path retries in full
                    if (sp.stage == 2 && sp.quick && nconv < ne && sp.iter >= 2 && sp.nsteps >= 80 + ctx->cheb_quick && evals[ne - 1] > -1.0) { sp.stage = 6; return; }
                    if (nconv >= ne || sp.iter >= sp.maxit) {
                        sp.nrestart += sp.iter + 1;
                        sp.evals = evals;
                        sp.Y = Y;
                        const bool conv = nconv >= ne;
                        if (sp.stage == 1) {
                            if (!accel || (conv && !sp.quick)) {
                                if (!conv) { sp.stage = 6; return; }  // plain mode ran out of restarts: let the single-point path report it
                                sp.final_evals.assign(evals.begin(), evals.begin() + nev);
                                sp.final_y.assign(Y.begin(), Y.begin() + nv);
                                sp.stage = 4;
                                return;
                            }
                            // -> stage 2: place the cut (bh_lanczos), start from the wanted Ritz vectors
                            const double th0 = evals[0], thn = sp.quick ? evals[0] : evals[nev - 1];
                            double hi = sp.hi;
                            hi += 1e-9 * (hi - sp.lo) + 1e-12;
                            const double cut = std::max(thn + ctx->cheb_margin * (thn - th0), th0 + ctx->cheb_frac * (hi - th0));
                            if (!(cut < th0 + 0.8 * (hi - th0))) { sp.stage = 6; return; }
                            sp.cut = cut;
                            sp.c = 0.5 * (hi + cut);
                            sp.e = 0.5 * (hi - cut);
                            sp.final_y.assign(nv, 0.0);  // coefficients of the start vector
                            const int take = sp.quick ? 1 : nev;
                            for (int l = 0; l < take; ++l)
                                for (int r = 0; r < nv; ++r) sp.final_y[r] += Y[r + (size_t)l * nv];
                            sp.stage = -2;  // start vector pending
                            return;
                        }
                        // stage 2 finished
                        if (!conv) { sp.stage = 6; return; }
                        sp.stage = 3;
                        return;
                    }
                    ++sp.iter;
                    int knew = ne;
                    for (int l = ne; l < nv; ++l)
                        if (std::fabs(Y[(nv - 1) + (size_t)l * nv]) < near0) ++knew;
                    knew += std::min(nconv, (nv - knew) / 2);
                    if (knew == 1 && nv >= 6) knew = nv / 2;
                    else if (knew == 1 && nv > 2) knew = 2;
                    if (knew > nv - 1) knew = nv - 1;
                    sp.theta.assign(evals.begin(), evals.begin() + knew);
                    sp.coup.resize(knew);
                    for (int l = 0; l < knew; ++l) sp.coup[l] = beta_last * Y[(nv - 1) + (size_t)l * nv];
                    sp.Y = Y;
                    sp.compress_k = knew;
                    sp.from = knew;
                });
                t_host += secs_since(t_h);
            }
            // ---- start vectors of stage 2 ----
            for (int p = 0; p < np; ++p) nlin += (pts[p].stage == -2);
            if (nlin) {
                for (int p = 0; p < np; ++p) {
                    SmallPoint& sp = pts[p];
                    desc[p].active = (sp.stage == -2);
                    if (!desc[p].active) continue;
                    desc[p].ncv = sp.ncv;
                    std::copy(sp.final_y.begin(), sp.final_y.end(), h_Y.begin() + (size_t)p * SM_MAX_NCV * SM_MAX_NCV);
                }
                BH_H2D(ctx, ws->d_desc, desc.data(), sizeof(SmallDesc) * np);
                BH_H2D(ctx, ws->Y, h_Y.data(), sizeof(double) * h_Y.size());
                k_small_lincomb<<<np, SM_THREADS, 0, st>>>(D, ws->d_desc, P, 0);
                BH_LAUNCHED(ctx);
                for (int p = 0; p < np; ++p) {
                    SmallPoint& sp = pts[p];
                    if (sp.stage != -2) continue;
                    sp.stage = 2;
                    sp.nev = nev; sp.ncv = ncv;
                    sp.maxit = sp.quick ? std::min(maxit, 80) : maxit;
                    sp.from = 0; sp.iter = 0; sp.init = 1; sp.compress_k = 0;
                    sp.theta.clear(); sp.coup.clear();
                }
            }
            // ---- Rayleigh-Ritz of H for the points whose filtered iteration has converged: ONE launch for all of them, once no
            // point is iterating any more (a launch per restart round kept a handful of SMs busy for 15-20 ms each: 0.2 s of the
            // 1.0 s of a C2 sweep) ----
            bool iterating = false;
            for (int p = 0; p < np; ++p) iterating = iterating || pts[p].stage == 1 || pts[p].stage == 2 || pts[p].stage == -2;
            for (int p = 0; p < np; ++p) nrr += (pts[p].stage == 3);
            if (nrr && !iterating) {
                for (int p = 0; p < np; ++p) {
                    desc[p].active = (pts[p].stage == 3);
                    desc[p].ncv = ncv;
                }
                BH_H2D(ctx, ws->d_desc, desc.data(), sizeof(SmallDesc) * np);
                {
                    BhProfScope prof(ctx, BH_PROF_SMALL, 0.0);
                    rrk<<<np, SM_THREADS, 0, st>>>(ctx->d_tab, D, ctx->d_states, ctx->d_dU, ws->d_desc, P);
                }
                BH_LAUNCHED(ctx);
                BH_CUDA(ctx, cudaGetLastError());
                BH_D2H(ctx, h_Y.data(), ws->Y, sizeof(double) * h_Y.size());
                BH_CUDA(ctx, cudaStreamSynchronize(st));
                parallel_for(np, [&](int p) {
                    SmallPoint& sp = pts[p];
                    if (sp.stage != 3) return;
                    sp.nmatvec += ncv;
                    std::vector<double> Mh((size_t)ncv * ncv), evH, Z;
                    const double* src = h_Y.data() + (size_t)p * SM_MAX_NCV * SM_MAX_NCV;
                    for (int a = 0; a < ncv; ++a)
                        for (int b = 0; b < ncv; ++b) Mh[a + (size_t)b * ncv] = 0.5 * (src[a + (size_t)b * ncv] + src[b + (size_t)a * ncv]);
                    bh_sym_eig(ncv, Mh, evH, Z);
                    const bool ok = sp.quick ? (evH[nev - 1] < sp.cut - 0.02 * (sp.cut - evH[0])) : (evH[nev - 1] < sp.cut);
                    if (!ok) { sp.stage = 6; return; }
                    sp.final_evals.assign(evH.begin(), evH.begin() + nev);
                    sp.final_y.assign(Z.begin(), Z.begin() + ncv);
                    sp.stage = 4;
                });
            }
            bool busy = false;
            for (int p = 0; p < np; ++p) busy = busy || (pts[p].stage >= 1 && pts[p].stage <= 3) || pts[p].stage == -2;
            if (!busy) break;
        }

        // ---- ground-state Ritz vectors, SPDM, output columns ----
        const auto t_f = now();
        int nfin = 0;
        for (int p = 0; p < np; ++p) {
            SmallPoint& sp = pts[p];
            desc[p].active = (sp.stage == 4);
            nfin += desc[p].active;
            if (!desc[p].active) continue;
            desc[p].ncv = (int)sp.final_y.size();
            std::copy(sp.final_y.begin(), sp.final_y.end(), h_Y.begin() + (size_t)p * SM_MAX_NCV * SM_MAX_NCV);
        }
        std::vector<double> rho((size_t)np * m * m, 0.0);
        if (nfin) {
            BH_H2D(ctx, ws->d_desc, desc.data(), sizeof(SmallDesc) * np);
            BH_H2D(ctx, ws->Y, h_Y.data(), sizeof(double) * h_Y.size());
            k_small_lincomb<<<np, SM_THREADS, 0, st>>>(D, ws->d_desc, P, 1);
            BH_LAUNCHED(ctx);
            BH_TRY(bh_spdm_batch_dev(ctx, ws->vec, 5 * ld, np, nb_eigen, ws->rho_part, ws->rho, rho.data()));
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        for (int p = 0; p < np; ++p) {
            SmallPoint& sp = pts[p];
            if (sp.stage != 4) {
                later.push_back(p0 + p);
                continue;
            }
            std::vector<double> ratios(nb_eigen - 2);
            bh_gap_ratios(sp.final_evals.data(), nb_eigen, ratios.data());
            double g = 0;
            for (double r : ratios) g += r;
            double* o = out3 + 3 * (p0 + p);
            o[0] = ratios.empty() ? 0.0 : g / (double)ratios.size();
            bh_condensate_fraction(m, rho.data() + (size_t)p * m * m, &o[1]);
            bh_coherence(m, rho.data() + (size_t)p * m * m, &o[2]);
            if (infos) {
                infos[p0 + p].nconv = nb_eigen;
                infos[p0 + p].nmatvec = sp.nmatvec;
                infos[p0 + p].nrestart = sp.nrestart;
                infos[p0 + p].nreorth = sp.nsteps;
                infos[p0 + p].seconds = secs / (double)npoints;
            }
        }
        t_final += secs_since(t_f);
    }
    if (verbose)
        fprintf(stderr, "[bh] small solver: %lld points, %d cycle launches, %zu to the single-point path; set-up %.3f s, waiting for the GPU %.3f s, "
                        "host Ritz problems %.3f s, final %.3f s, total %.3f s\n",
                (long long)npoints, ncycles_total, later.size(), t_setup, t_gpu_wait, t_host, t_final, secs_since(t_start));
    // ---- stragglers: breakdowns (J = 0) and failed safety checks go through the ordinary path ----
    for (int64_t p : later)
        BH_TRY(bh_point(ctx, cJ[p], cU[p], cmu[p], nb_eigen, BH_HV_MATRIX_FREE, out3 + 3 * p, nullptr, nullptr, infos ? infos + p : nullptr));
    return BH_OK;
}
