// observables.cu -- K7: ground-state observables as device reductions, and the per-point driver.
//
// Replaces Analysis::SPDM / braket / coherence / gap_ratios and the body of the sweep loop
// (reference src/analysis.cpp:311-337,433-454,497-594).  The reference makes m(m+1)/2 passes over the D
// basis states with a tag + binary search per term; here one pass per source site j accumulates every
// <a_i^+ a_j>, i <= j, with the O(1) incremental rank.  Kept quirks: the amplitude is
// sqrt((n_i + 1)(n_j - 1)) for i != j, terms with |phi| <= eps are skipped, and rho is
// divided by the number of eigenvector columns (src/analysis.cpp:527).
#include <algorithm>
#include <cmath>
#include <limits>

#include "device_utils.cuh"

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

#define SPDM_THREADS 256

template <int M>
__global__ void __launch_bounds__(SPDM_THREADS)
k_spdm(const BhTables* __restrict__ gtab, int64_t row0, int64_t D, const uint64_t* __restrict__ states,
       const double* __restrict__ phi /* full vector, indexed by global LEX rank */, double* __restrict__ part /* [gridDim.x][M][M] */,
       int64_t phi_stride /* blockIdx.z selects one of several vectors (batched form) */)
{
    phi += (int64_t)blockIdx.z * phi_stride;
    part += (int64_t)blockIdx.z * gridDim.x * M * M;
    __shared__ BhTables t;
    __shared__ double scratch[32];
    bh_stage_tables(&t, gtab);
    const int j = blockIdx.y;  // source site
    const double eps = 2.220446049250313e-16;
    double acc[M];
#pragma unroll
    for (int i = 0; i < M; ++i) acc[i] = 0.0;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < D; l += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = row0 + l;
        const double pk = phi[k];
        if (!(fabs(pk) > eps)) continue;
        const uint64_t s = states[l];
        const int nj = bh_occ(s, j);
        if (nj < 1) continue;
        int dn[M], up[M];
        bh_rank_prefix<M>(t, s, dn, up);
        int dnj = 0;
#pragma unroll
        for (int q = 0; q < M; ++q)
            if (q == j) dnj = dn[q];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            if (i > j) continue;
            if (i == j) {
                acc[i] += pk * pk * t.sq[nj * nj];  // sqrt(n_j * n_j) = n_j
            } else {
                const int tgt = (int)k + dnj - dn[i];
                const double pt = __ldg(phi + tgt);
                if (fabs(pt) > eps) acc[i] += pk * pt * t.sq[(bh_occ(s, i) + 1) * (nj - 1)];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
        const double v = bh_block_sum(acc[i], scratch);
        if (threadIdx.x == 0) part[((int64_t)blockIdx.x * M + j) * M + i] = v;
    }
}

__global__ void k_spdm_reduce(int m, int nb, const double* __restrict__ part, double inv_cols, double* __restrict__ rho)
{
    part += (int64_t)blockIdx.x * nb * m * m;
    rho += (int64_t)blockIdx.x * m * m;
    const int i = threadIdx.x % m, j = threadIdx.x / m;
    if (j >= m || i > j) return;
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += part[((int64_t)b * m + j) * m + i];
    t *= inv_cols;
    rho[i + j * m] = t;  // column-major, upper triangle computed, lower mirrored (src/analysis.cpp:509-511)
    rho[j + i * m] = t;
}

typedef void (*spdm_fn)(const BhTables*, int64_t, int64_t, const uint64_t*, const double*, double*, int64_t);
static spdm_fn spdm_kernel(int m)
{
    switch (m) {
        case 1: return k_spdm<1>;
        case 2: return k_spdm<2>;
        case 3: return k_spdm<3>;
        case 4: return k_spdm<4>;
        case 5: return k_spdm<5>;
        case 6: return k_spdm<6>;
        case 7: return k_spdm<7>;
        case 8: return k_spdm<8>;
        case 9: return k_spdm<9>;
        case 10: return k_spdm<10>;
        case 11: return k_spdm<11>;
        case 12: return k_spdm<12>;
        case 13: return k_spdm<13>;
        case 14: return k_spdm<14>;
        case 15: return k_spdm<15>;
        case 16: return k_spdm<16>;
    }
    return nullptr;
}

int bh_spdm_dev(bh_ctx* ctx, const double* phi_dev, int ncols, double* rho_host)
{
    const int m = ctx->m;
    BH_TRY(bh_ensure_workspace(ctx, 0));
    const int64_t nloc = ctx->nloc;
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks(nloc, SPDM_THREADS), (int64_t)ctx->sm_count * 2));
    // scratch lives in the context (no cudaMalloc / cudaFree on the per-point path: both synchronise the device)
    const size_t need = sizeof(double) * ((size_t)gx * m * m + (size_t)m * m);
    if (ctx->spdm_scratch_bytes < need) {
        if (ctx->d_spdm_scratch) cudaFree(ctx->d_spdm_scratch);
        ctx->d_spdm_scratch = nullptr;
        BH_CUDA(ctx, cudaMalloc(&ctx->d_spdm_scratch, need));
        ctx->spdm_scratch_bytes = need;
    }
    double* d_part = ctx->d_spdm_scratch;
    double* d_rho = d_part + (size_t)gx * m * m;
    const bool dist = ctx->partitioned && ctx->world > 1;
    const double* phi_full = phi_dev;
    if (dist) {  // phi_dev is the local slice: exchange it once, reduce the local partial matrices afterwards
        BH_TRY(bh_dist_allgather(ctx, phi_dev, ctx->d_xfull, ctx->ld));
        phi_full = ctx->d_xfull;
    }
    dim3 grid(gx, m);
    {
        BhProfScope prof(ctx, BH_PROF_SPDM, 16.0 * (double)nloc * m);  // phi and the packed states, once per source site
        spdm_kernel(m)<<<grid, SPDM_THREADS, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, nloc, ctx->d_states, phi_full, d_part, 0);
        k_spdm_reduce<<<1, m * m, 0, ctx->stream>>>(m, gx, d_part, 1.0 / (double)ncols, d_rho);
    }
    if (dist) BH_TRY(bh_dist_allreduce_sum(ctx, d_rho, m * m));
    ctx->launches += 2;
    BH_CUDA(ctx, cudaGetLastError());
    BH_D2H(ctx, rho_host, d_rho, sizeof(double) * m * m);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BH_OK;
}

// rho of npts vectors phi + p * stride (unpartitioned context) with one launch pair; d_part: npts * 8 * m * m doubles
int bh_spdm_batch_dev(bh_ctx* ctx, const double* phi, int64_t stride, int npts, int ncols, double* d_part, double* d_rho, double* rho_host)
{
    const int m = ctx->m;
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks(ctx->nloc, SPDM_THREADS), 8));
    dim3 grid(gx, m, npts);
    {
        BhProfScope prof(ctx, BH_PROF_SPDM, 16.0 * (double)ctx->nloc * m * npts);
        spdm_kernel(m)<<<grid, SPDM_THREADS, 0, ctx->stream>>>(ctx->d_tab, 0, ctx->nloc, ctx->d_states, phi, d_part, stride);
        k_spdm_reduce<<<npts, m * m, 0, ctx->stream>>>(m, gx, d_part, 1.0 / (double)ncols, d_rho);
    }
    ctx->launches += 2;
    BH_CUDA(ctx, cudaGetLastError());
    BH_D2H(ctx, rho_host, d_rho, sizeof(double) * (size_t)npts * m * m);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BH_OK;
}

extern "C" int bh_spdm(bh_ctx* ctx, int order, const double* phi, int ncols, double* rho)
{
    if (!ctx || !ctx->D || ctx->user_matrix || ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_spdm: call bh_setup first (host vectors are not accepted on a row-partitioned context)");
    if (!phi || !rho || ncols < 1 || order < 0 || order > 2) return bh_fail(ctx, BH_ERR_ARG, "bh_spdm: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    BH_TRY(bh_ensure_staging(ctx));
    BH_H2D(ctx, ctx->d_y, phi, sizeof(double) * ctx->D);
    BH_TRY(bh_permute_vec(ctx, order, false, ctx->d_y, ctx->d_x));
    return bh_spdm_dev(ctx, ctx->d_x, ncols, rho);
}

// ---- host scalars ----
extern "C" int bh_gap_ratios(const double* evals, int nb_eigen, double* ratios)
{
    if (!evals || !ratios || nb_eigen < 3) return BH_ERR_ARG;
    std::vector<double> s(evals, evals + nb_eigen);
    std::sort(s.begin(), s.end());
    for (int i = 1; i < nb_eigen - 1; ++i) {
        const double up = s[i + 1] - s[i], dn = s[i] - s[i - 1];
        const double lo = std::min(up, dn), hi = std::max(up, dn);
        ratios[i - 1] = (hi != 0) ? lo / hi : 0.0;
    }
    return BH_OK;
}

extern "C" int bh_condensate_fraction(int m, const double* rho, double* out)
{
    if (!rho || !out || m < 1) return BH_ERR_ARG;
    std::vector<double> a(rho, rho + (size_t)m * m), ev, vec;
    double tr = 0;
    for (int i = 0; i < m; ++i) tr += rho[i + (size_t)i * m];
    bh_sym_eig(m, a, ev, vec);
    double best = ev[0];
    for (int i = 1; i < m; ++i)
        if (std::fabs(best) < std::fabs(ev[i])) best = ev[i];
    *out = std::fabs(best / tr);
    return BH_OK;
}

extern "C" int bh_coherence(int m, const double* rho, double* out)
{
    if (!rho || !out || m < 1) return BH_ERR_ARG;
    double all = 0, diag = 0;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            const double p = rho[i + (size_t)j * m] * rho[j + (size_t)i * m];
            all += p;
            if (i == j) diag += p;
        }
    *out = (all - diag) / all;
    return BH_OK;
}

// ---- finite-temperature branch (src/analysis.cpp:456-494) ----
extern "C" int bh_thermal_weights(const double* evals, int nb_eigen, double temperature, double* weights)
{
    if (!evals || !weights || nb_eigen < 1 || !(temperature > 0.0)) return BH_ERR_ARG;
    // :481  normalized = eigenvalues / max(eigenvalues);  :484  Z = sum exp(-beta normalized);  :486  w = exp(-beta normalized) / Z
    double mx = evals[0];
    for (int i = 1; i < nb_eigen; ++i) mx = std::max(mx, evals[i]);
    const double beta = 1.0 / temperature;
    double Z = 0.0;
    for (int i = 0; i < nb_eigen; ++i) Z += std::exp(-beta * (evals[i] / mx));
    for (int i = 0; i < nb_eigen; ++i) weights[i] = std::exp(-beta * (evals[i] / mx)) / Z;
    return BH_OK;
}

// rho[i + j D] = sum_k w[k] U[i + k D] U[j + k D]  (16 x 16 tiles, the eigenvector slabs staged in shared memory)
__global__ void __launch_bounds__(256)
k_density_matrix(int64_t D, int nk, const double* __restrict__ U, const double* __restrict__ w, double* __restrict__ rho)
{
    __shared__ double ui[16][65], uj[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i0 = (int64_t)blockIdx.x * 16, j0 = (int64_t)blockIdx.y * 16;
    for (int idx = threadIdx.x; idx < 16 * nk; idx += 256) {
        const int r = idx & 15, k = idx >> 4;
        ui[r][k] = (i0 + r < D) ? U[i0 + r + (int64_t)k * D] * w[k] : 0.0;
        uj[r][k] = (j0 + r < D) ? U[j0 + r + (int64_t)k * D] : 0.0;
    }
    __syncthreads();
    double acc = 0.0;
    for (int k = 0; k < nk; ++k) acc = fma(ui[tx][k], uj[ty][k], acc);
    if (i0 + tx < D && j0 + ty < D) rho[(i0 + tx) + (j0 + ty) * D] = acc;
}

extern "C" int bh_density_matrix(bh_ctx* ctx, int64_t D, int nb_eigen, const double* evals, const double* evecs, double temperature,
                                 double* rho)
{
    if (!ctx || !evals || !evecs || !rho || D < 1 || nb_eigen < 1 || nb_eigen > 64) return bh_fail(ctx, BH_ERR_ARG, "bh_density_matrix: bad argument");
    if (D > 46000) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "bh_density_matrix: the dense D x D matrix of the reference is limited to D <= 46000 here");
    std::vector<double> w(nb_eigen);
    if (bh_thermal_weights(evals, nb_eigen, temperature, w.data()) != BH_OK) return bh_fail(ctx, BH_ERR_ARG, "bh_density_matrix: temperature must be > 0");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dU = nullptr, *dw = nullptr, *drho = nullptr;
    BH_CUDA(ctx, cudaMalloc(&dU, sizeof(double) * (size_t)D * nb_eigen));
    BH_CUDA(ctx, cudaMalloc(&dw, sizeof(double) * nb_eigen));
    BH_CUDA(ctx, cudaMalloc(&drho, sizeof(double) * (size_t)D * D));
    BH_H2D(ctx, dU, evecs, sizeof(double) * (size_t)D * nb_eigen);
    BH_H2D(ctx, dw, w.data(), sizeof(double) * nb_eigen);
    dim3 grid((unsigned)((D + 15) / 16), (unsigned)((D + 15) / 16));
    k_density_matrix<<<grid, 256, 0, ctx->stream>>>(D, nb_eigen, dU, dw, drho);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    BH_D2H(ctx, rho, drho, sizeof(double) * (size_t)D * D);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(dU); cudaFree(dw); cudaFree(drho);
    return BH_OK;
}

// ---- one grid point / a shard of grid points ----
extern "C" int bh_point(bh_ctx* ctx, double cJ, double cU, double cmu, int nb_eigen, int kernel, double* out3,
                        double* evals, double* rho, bh_eigs_info* info)
{
    if (!ctx || !ctx->D || ctx->user_matrix) return bh_fail(ctx, BH_ERR_STATE, "bh_point: call bh_setup first");
    if (!out3 || nb_eigen < 3) return bh_fail(ctx, BH_ERR_ARG, "bh_point: bad argument");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    BhSolve s;
    // Op::IRLM_eigen: nev = nb_eigen, ncv = 2 nb_eigen + 1, Spectra defaults tol 1e-10, maxit 1000
    const int rc = bh_lanczos(ctx, cJ, cU, cmu, nb_eigen, 2 * nb_eigen + 1, 1e-10, 1000, kernel, &s);
    if (info) *info = s.info;
    if (rc != BH_OK) return rc;
    std::vector<double> ratios(nb_eigen - 2);
    bh_gap_ratios(s.evals.data(), nb_eigen, ratios.data());
    double g = 0;
    for (double r : ratios) g += r;
    out3[0] = ratios.empty() ? 0.0 : g / (double)ratios.size();
    BH_TRY(bh_ensure_staging(ctx));
    BH_TRY(bh_ritz_vector(ctx, s, 0, ctx->d_x));
    std::vector<double> r((size_t)ctx->m * ctx->m);
    BH_TRY(bh_spdm_dev(ctx, ctx->d_x, nb_eigen, r.data()));
    bh_condensate_fraction(ctx->m, r.data(), &out3[1]);
    bh_coherence(ctx->m, r.data(), &out3[2]);
    if (evals) std::copy(s.evals.begin(), s.evals.end(), evals);
    if (rho) std::copy(r.begin(), r.end(), rho);
    return BH_OK;
}

extern "C" int bh_points(bh_ctx* ctx, const double* cJ, const double* cU, const double* cmu, int64_t npoints, int nb_eigen,
                         int kernel, double* out3, bh_eigs_info* infos)
{
    if (!ctx || !cJ || !cU || !cmu || !out3 || npoints < 0) return bh_fail(ctx, BH_ERR_ARG, "bh_points: bad argument");
    // many points of a small system (small.cu): one CTA per grid point, a whole restart cycle per launch
    if (ctx->D && bh_small_supported(ctx, kernel, npoints, nb_eigen)) return bh_points_small(ctx, npoints, cJ, cU, cmu, nb_eigen, out3, infos);
    // lockstep batching (batch.cu): ctx->batch solves run together and share their H.v launches; same results point by point
    if (ctx->batch >= 2 && npoints >= 2 && ctx->D && nb_eigen >= 3 && bh_batch_supported(ctx, kernel))
        return bh_points_lockstep(ctx, (int)std::min<int64_t>(ctx->batch, npoints), npoints, cJ, cU, cmu, nb_eigen, kernel, out3, infos);
    for (int64_t p = 0; p < npoints; ++p)
        BH_TRY(bh_point(ctx, cJ[p], cU[p], cmu[p], nb_eigen, kernel, out3 + 3 * p, nullptr, nullptr, infos ? infos + p : nullptr));
    return BH_OK;
}
