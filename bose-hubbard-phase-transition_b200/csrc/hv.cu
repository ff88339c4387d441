// hv.cu -- K3 (stored-matrix H.v) and K4 (matrix-free H.v).
//
// Replaces Spectra::SparseGenMatProd::perform_op -> Eigen's serial CSC scatter product
// (reference external/spectra/include/Spectra/MatOp/SparseGenMatProd.h:81-86,
//  include/Eigen/src/SparseCore/SparseDenseProduct.h:86-107).
//
// K3  streams the CSR arrays of the materialised H once per product, fully coalesced: a warp owns 32
//     consecutive rows, its lanes stride over the contiguous entry range of those rows (one coalesced
//     128 B / 256 B request per warp instruction for col / val), gather x through L2, park the products
//     in shared memory and then each lane adds up its own row in column order (deterministic, the same
//     summation order as the reference's column-ordered scatter).  Algorithmic bytes: 12*nnz + 4*(D+1) + 16*D.
// K4  stores no matrix: thread k re-derives row k from the packed state (8 B) with the O(1) incremental
//     rank, so the HBM traffic is x, y, the packed states and the U-diagonal only.
#include "device_utils.cuh"

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------
#define HV_WARPS 8

__global__ void __launch_bounds__(HV_WARPS * 32)
k_hv_csr(int64_t D, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
         const double* __restrict__ x, double* __restrict__ y, int max_row)
{
    extern __shared__ double sprod[];
    const int lane = threadIdx.x & 31;
    double* prod = sprod + (size_t)(threadIdx.x >> 5) * 32 * max_row;
    const int64_t ntiles = (D + 31) >> 5;
    const int64_t wstride = (int64_t)gridDim.x * HV_WARPS;
    for (int64_t tile = (int64_t)blockIdx.x * HV_WARPS + (threadIdx.x >> 5); tile < ntiles; tile += wstride) {
        const int64_t r = (tile << 5) + lane;
        int a = 0, b = 0;
        if (r < D) {
            a = __ldg(rowptr + r);
            b = __ldg(rowptr + r + 1);
        }
        const int e0 = __shfl_sync(0xffffffffu, a, 0);
        const int nvalid = (int)min((int64_t)32, D - (tile << 5));
        const int e1 = __shfl_sync(0xffffffffu, b, nvalid - 1);
        // phase 1: products, entry-parallel and coalesced; 4 independent requests in flight per lane
        int e = e0 + lane;
        for (; e + 96 < e1; e += 128) {
            const int c0 = __ldg(col + e), c1 = __ldg(col + e + 32), c2 = __ldg(col + e + 64), c3 = __ldg(col + e + 96);
            const double v0 = __ldg(val + e), v1 = __ldg(val + e + 32), v2 = __ldg(val + e + 64), v3 = __ldg(val + e + 96);
            const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
            prod[e - e0] = v0 * x0;
            prod[e - e0 + 32] = v1 * x1;
            prod[e - e0 + 64] = v2 * x2;
            prod[e - e0 + 96] = v3 * x3;
        }
        for (; e < e1; e += 32) prod[e - e0] = __ldg(val + e) * __ldg(x + __ldg(col + e));
        __syncwarp();
        // phase 2: every lane sums its own row in column order
        if (r < D) {
            double acc = 0.0;
            for (int q = a - e0; q < b - e0; ++q) acc += prod[q];
            y[r] = acc;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256)
k_hv_free(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states,
          const double* __restrict__ dU, double cJ, double cU, double cmu, const double* __restrict__ x,
          double* __restrict__ y)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const double shift = __dmul_rn(-(double)t.n, cmu);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < D; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t s = states[k];
        int dn[M], up[M];
        bh_rank_prefix<M>(t, s, dn, up);
        double acc = 0.0;
#pragma unroll
        for (int src = 0; src < M; ++src) {
            const int ns = bh_occ(s, src);
            if (ns == 0) continue;
#pragma unroll
            for (int dst = 0; dst < M; ++dst) {
                if (dst == src) continue;
                const int w = t.w[dst][src];  // block-uniform
                if (w == 0) continue;
                const int tgt = (int)k + (dst < src ? dn[src] - dn[dst] : up[dst] - up[src]);
                acc += (double)w * t.sq[(bh_occ(s, dst) + 1) * ns] * __ldg(x + tgt);
            }
        }
        const double diag = __dadd_rn(__dmul_rn(dU[k], cU), shift);
        y[k] = diag * x[k] - cJ * acc;
    }
}

typedef void (*hv_free_fn)(const BhTables*, int64_t, const uint64_t*, const double*, double, double, double,
                           const double*, double*);

static hv_free_fn hv_free_kernel(int m)
{
    switch (m) {
        case 1: return k_hv_free<1>;
        case 2: return k_hv_free<2>;
        case 3: return k_hv_free<3>;
        case 4: return k_hv_free<4>;
        case 5: return k_hv_free<5>;
        case 6: return k_hv_free<6>;
        case 7: return k_hv_free<7>;
        case 8: return k_hv_free<8>;
        case 9: return k_hv_free<9>;
        case 10: return k_hv_free<10>;
        case 11: return k_hv_free<11>;
        case 12: return k_hv_free<12>;
        case 13: return k_hv_free<13>;
        case 14: return k_hv_free<14>;
        case 15: return k_hv_free<15>;
        case 16: return k_hv_free<16>;
    }
    return nullptr;
}

int bh_launch_hv(bh_ctx* ctx, double cJ, double cU, double cmu, int kernel, const double* x, double* y, double)
{
    const int64_t D = ctx->D;
    if (kernel == BH_HV_STORED) {
        BH_TRY(bh_materialise_H(ctx, cJ, cU, cmu));
        const size_t smem = (size_t)HV_WARPS * 32 * ctx->max_row * sizeof(double);
        static thread_local size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            BH_CUDA(ctx, cudaFuncSetAttribute(k_hv_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
        const int64_t ntiles = (D + 31) / 32;
        int grid = (int)std::min<int64_t>((ntiles + HV_WARPS - 1) / HV_WARPS, (int64_t)ctx->sm_count * per_sm);
        k_hv_csr<<<grid, HV_WARPS * 32, smem, ctx->stream>>>(D, ctx->d_rowptr, ctx->d_col, ctx->d_valH, x, y,
                                                             ctx->max_row);
        BH_LAUNCHED(ctx);
    } else if (kernel == BH_HV_MATRIX_FREE) {
        hv_free_fn fn = hv_free_kernel(ctx->m);
        int grid = (int)std::min<int64_t>(nblocks(D, 256), (int64_t)ctx->sm_count * 8);
        fn<<<grid, 256, 0, ctx->stream>>>(ctx->d_tab, D, ctx->d_states, ctx->d_dU, cJ, cU, cmu, x, y);
        BH_LAUNCHED(ctx);
    } else {
        return bh_fail(ctx, BH_ERR_ARG, "unknown H.v kernel");
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
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    BH_TRY(bh_ensure_staging(ctx));
    BH_TRY(bh_ensure_workspace(ctx, 0));
    const size_t bytes = sizeof(double) * ctx->D;
    BH_CUDA(ctx, cudaMemcpyAsync(ctx->d_y, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    BH_TRY(bh_permute_vec(ctx, order, false, ctx->d_y, ctx->d_x));   // x_lex
    BH_TRY(bh_launch_hv(ctx, cJ, cU, cmu, kernel, ctx->d_x, ctx->d_w));
    BH_TRY(bh_permute_vec(ctx, order, true, ctx->d_w, ctx->d_y));    // y in `order`
    BH_CUDA(ctx, cudaMemcpyAsync(y, ctx->d_y, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BH_OK;
}

extern "C" int bh_hv_algorithmic_bytes(bh_ctx* ctx, int kernel, int64_t* bytes)
{
    if (!ctx || !ctx->D || !bytes) return bh_fail(ctx, BH_ERR_STATE, "bh_hv_algorithmic_bytes: call bh_setup first");
    if (kernel == BH_HV_STORED)
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
__global__ void k_lcg_fill(int64_t n, double* __restrict__ out)
{
    const unsigned long long M = 2147483647ull;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t start = c * LCG_CHUNK;
    if (start >= n) return;
    unsigned long long seed = 1, base = 16807ull;
    for (unsigned long long p = (unsigned long long)start; p; p >>= 1) {
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
    k_lcg_fill<<<nblocks((count + LCG_CHUNK - 1) / LCG_CHUNK, 128), 128, 0, ctx->stream>>>(count, x_dev);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}
