// ctx_basis.cu -- context management and K1: Fock-basis enumeration / ranking / tags / orderings.
//
// Replaces BH::binomial/dimension/init_lexicographic/calculate_tags/sort_basis/search_tag/fixed_set_basis
// (reference src/hamiltonian.cpp:39-149).  The basis is never enumerated sequentially: thread k unranks
// state k from the binomial table (descending-lexicographic order = the reference's enumeration order),
// tags are evaluated with un-fused multiply/add in site order so they are bit-identical to the reference,
// and the tag ordering is a device radix sort on the raw bit patterns of the (positive) tags.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "device_utils.cuh"

static std::string g_create_error;
static std::mutex g_create_mutex;

int bh_fail(bh_ctx* ctx, int code, const std::string& msg)
{
    if (ctx)
        ctx->err = msg;
    else {
        std::lock_guard<std::mutex> lk(g_create_mutex);
        g_create_error = msg;
    }
    return code;
}

extern "C" const char* bh_last_error(const bh_ctx* ctx)
{
    if (ctx) return ctx->err.c_str();
    return g_create_error.c_str();
}

extern "C" int bh_ctx_create(int device, bh_ctx** out)
{
    if (!out) return bh_fail(nullptr, BH_ERR_ARG, "bh_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return bh_fail(nullptr, BH_ERR_CUDA,
                       std::string("bh_ctx_create: no CUDA device (") + cudaGetErrorString(e) +
                           "); this library has no CPU path");
    if (device < 0 || device >= count) return bh_fail(nullptr, BH_ERR_ARG, "bh_ctx_create: bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return bh_fail(nullptr, BH_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    bh_ctx* ctx = new bh_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return bh_fail(nullptr, BH_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    ctx->own_stream = true;
    if (const char* v = getenv("BH_BATCH")) ctx->batch = std::min(4, std::max(1, atoi(v)));
    if (const char* v = getenv("BH_BATCH_PLAIN")) ctx->batch_plain = atoi(v);
    if (const char* v = getenv("BH_SELL_SIGMA")) ctx->sell_sigma = std::min(1024, std::max(32, atoi(v) / 32 * 32));
    if (const char* v = getenv("BH_COOP")) ctx->coop = atoi(v);
    if (const char* v = getenv("BH_COOP_FUSED")) ctx->coop_fused = atoi(v);
    if (const char* v = getenv("BH_RR_GRAM")) ctx->rr_gram = atoi(v);
    if (const char* v = getenv("BH_COMPRESS_TILED")) ctx->compress_tiled = atoi(v);
    if (const char* v = getenv("BH_CHEB_DEGREE")) ctx->cheb_degree = std::max(1, atoi(v));
    if (const char* v = getenv("BH_CHEB_PRE")) ctx->cheb_pre = std::max(1, atoi(v));
    if (const char* v = getenv("BH_CHEB_QUICK")) ctx->cheb_quick = std::max(0, atoi(v));
    if (const char* v = getenv("BH_CHEB_MARGIN")) ctx->cheb_margin = atof(v);
    if (const char* v = getenv("BH_CHEB_FRAC")) ctx->cheb_frac = atof(v);
    if (const char* v = getenv("BH_REORTH_BLOCK")) { ctx->reorth_block = std::min(BH_MAX_NCV, std::max(1, atoi(v))); ctx->reorth_block_forced = true; }
    *out = ctx;
    return BH_OK;
}

static void free_dev(void* p)
{
    if (p) cudaFree(p);
}

void bh_release_workspace(bh_ctx* ctx)
{
    if (!ctx->parent) bh_small_release(ctx);  // lockstep children alias the pointer of their parent
    if (ctx->d_arena) bh_dist_arena_release(ctx);  // partitioned context: V, w, f, cheb live in one exported arena
    free_dev(ctx->d_V); ctx->d_V = nullptr;
    free_dev(ctx->d_w); ctx->d_w = nullptr;
    free_dev(ctx->d_f); ctx->d_f = nullptr;
    free_dev(ctx->d_scal); ctx->d_scal = nullptr;
    free_dev(ctx->d_part); ctx->d_part = nullptr;
    free_dev(ctx->d_counter); ctx->d_counter = nullptr;
    free_dev(ctx->d_small); ctx->d_small = nullptr;
    for (int q = 0; q < 3; ++q) { free_dev(ctx->d_cheb[q]); ctx->d_cheb[q] = nullptr; }
    free_dev(ctx->d_hv_block); ctx->d_hv_block = nullptr;
    ctx->hv_block_cols = 0;
    free_dev(ctx->d_gram_part); ctx->d_gram_part = nullptr;
    free_dev(ctx->d_spdm_scratch); ctx->d_spdm_scratch = nullptr;
    ctx->spdm_scratch_bytes = 0;
    free_dev(ctx->d_x); ctx->d_x = nullptr;
    free_dev(ctx->d_y); ctx->d_y = nullptr;
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr;
    ctx->h_pinned_bytes = 0;
    ctx->ws_ncv = 0;
}

int bh_release_system(bh_ctx* ctx)
{
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bh_batch_release(ctx);  // lockstep children alias the arrays freed below
    free_dev(ctx->d_tab); ctx->d_tab = nullptr;
    free_dev(ctx->d_states); ctx->d_states = nullptr;
    free_dev(ctx->d_dU); ctx->d_dU = nullptr;
    free_dev(ctx->d_rowptr); ctx->d_rowptr = nullptr;
    free_dev(ctx->d_col); ctx->d_col = nullptr;
    free_dev(ctx->d_valJ); ctx->d_valJ = nullptr;
    free_dev(ctx->d_diagpos); ctx->d_diagpos = nullptr;
    free_dev(ctx->d_valH); ctx->d_valH = nullptr;
    free_dev(ctx->d_sell_row); ctx->d_sell_row = nullptr;
    free_dev(ctx->d_sell_ptr); ctx->d_sell_ptr = nullptr;
    free_dev(ctx->d_sell_col); ctx->d_sell_col = nullptr;
    free_dev(ctx->d_sell_valJ); ctx->d_sell_valJ = nullptr;
    free_dev(ctx->d_sell_valH); ctx->d_sell_valH = nullptr;
    free_dev(ctx->d_sell_diag); ctx->d_sell_diag = nullptr;
    ctx->sell_valid = ctx->sell_partial_valid = false;
    ctx->sell_cJ = ctx->sell_cU = ctx->sell_cmu = 0.0;
    ctx->sell_nslices = ctx->sell_entries = 0;
    free_dev(ctx->d_tags); ctx->d_tags = nullptr;
    free_dev(ctx->d_perm_tag); ctx->d_perm_tag = nullptr;
    free_dev(ctx->d_inv_tag); ctx->d_inv_tag = nullptr;
    bh_release_workspace(ctx);
    ctx->valH_valid = false;
    ctx->m = ctx->n = 0;
    ctx->D = 0;
    ctx->user_matrix = false;
    bh_dist_release_halo(ctx);
    ctx->partitioned = false;
    ctx->row0 = ctx->nloc = 0;
    free_dev(ctx->d_xfull); ctx->d_xfull = nullptr;
    return BH_OK;
}

extern "C" int bh_ctx_destroy(bh_ctx* ctx)
{
    if (!ctx) return BH_OK;
    bh_release_system(ctx);
    if (ctx->nccl_comm) bh_dist_finalize(ctx);  // a context that joined a communicator leaves it here
    bh_ctx_profile_enable(ctx, 0);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return BH_OK;
}

// ---- per-kernel event timing ----
static cudaEvent_t prof_event(bh_ctx* root)
{
    if (!root->prof_free.empty()) {
        cudaEvent_t e = root->prof_free.back();
        root->prof_free.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void bh_prof_begin(bh_ctx* root, int cls, double bytes)
{
    BhProfRec r{prof_event(root), prof_event(root), cls, bytes};
    cudaEventRecord(r.a, root->stream);
    root->prof.push_back(r);
}

void bh_prof_end(bh_ctx* root) { cudaEventRecord(root->prof.back().b, root->stream); }

static void prof_collect(bh_ctx* ctx)
{
    if (ctx->prof.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (const BhProfRec& r : ctx->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.cls >= 0 && r.cls < BH_PROF_NCLASSES) {
            ctx->prof_ms[r.cls] += ms;
            ctx->prof_bytes[r.cls] += r.bytes;
            ctx->prof_n[r.cls]++;
        }
        ctx->prof_free.push_back(r.a);
        ctx->prof_free.push_back(r.b);
    }
    ctx->prof.clear();
}

extern "C" int bh_ctx_profile_enable(bh_ctx* ctx, int on)
{
    if (!ctx) return BH_ERR_ARG;
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    ctx->prof_on = on != 0;
    if (!on) {
        for (cudaEvent_t e : ctx->prof_free) cudaEventDestroy(e);
        ctx->prof_free.clear();
    }
    return BH_OK;
}

extern "C" int bh_ctx_profile_read(bh_ctx* ctx, int cls, int64_t* launches, double* total_ms, double* total_bytes)
{
    if (!ctx || cls < 0 || cls >= BH_PROF_NCLASSES) return bh_fail(ctx, BH_ERR_ARG, "bh_ctx_profile_read: bad class");
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    if (launches) *launches = ctx->prof_n[cls];
    if (total_ms) *total_ms = ctx->prof_ms[cls];
    if (total_bytes) *total_bytes = ctx->prof_bytes[cls];
    ctx->prof_n[cls] = 0;
    ctx->prof_ms[cls] = ctx->prof_bytes[cls] = 0.0;
    return BH_OK;
}

extern "C" int bh_ctx_set_stream(bh_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return BH_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (cuda_stream == nullptr) {
        if (!ctx->own_stream) {
            BH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
            ctx->own_stream = true;
        }
        return BH_OK;
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return BH_OK;
}

extern "C" int bh_ctx_set_batch(bh_ctx* ctx, int batch)
{
    if (!ctx) return BH_ERR_ARG;
    if (batch < 1 || batch > 4) return bh_fail(ctx, BH_ERR_ARG, "bh_ctx_set_batch: batch must be 1..4");
    ctx->batch = batch;
    return BH_OK;
}

extern "C" int64_t bh_ctx_launch_count(const bh_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int bh_ctx_transfer_bytes(const bh_ctx* ctx, int64_t* h2d, int64_t* d2h)
{
    if (!ctx) return BH_ERR_ARG;
    if (h2d) *h2d = ctx->h2d_bytes;
    if (d2h) *d2h = ctx->d2h_bytes;
    return BH_OK;
}

// ---- binomials (64-bit, exact) ----
static int64_t binom64(int n, int k)
{
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return (int64_t)r;
}

extern "C" int bh_dimension(int m, int n, int64_t* D)
{
    if (!D || m < 1 || n < 0 || m + n > 62) return BH_ERR_ARG;
    *D = binom64(m + n - 1, n);
    return BH_OK;
}

// ---- K1 kernels ----
// Thread k unranks state k: greedy descent through the table, site by site.
__global__ void k_unrank(const BhTables* __restrict__ gtab, int64_t row0, int64_t nloc, uint64_t* __restrict__ states,
                         double* __restrict__ dU)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // local row; global rank = row0 + k
    if (k >= nloc) return;
    int rem = (int)(row0 + k);
    int Rprev = t.n;
    uint64_t s = 0;
    int u = 0;
    for (int q = 0; q < t.m - 1; ++q) {
        int R = Rprev;
        while (t.f[q][R] > rem) --R;  // f[q][0] = 0 always terminates
        const int nq = Rprev - R;
        s |= (uint64_t)nq << (4 * q);
        u += nq * (nq + 1);
        rem -= t.f[q][R];
        Rprev = R;
    }
    s |= (uint64_t)Rprev << (4 * (t.m - 1));
    u += Rprev * (Rprev + 1);
    states[k] = s;
    dU[k] = (double)u;  // sum_i n_i (n_i + 1), src/hamiltonian.cpp:203-207
}

// tag_k = sum_i n_i * log(p_i), sequential in site order, separate multiply and add (src/hamiltonian.cpp:91-97)
__device__ __forceinline__ double bh_tag(const BhTables& t, uint64_t s)
{
    double tag = 0.0;
    for (int i = 0; i < t.m; ++i) tag = __dadd_rn(tag, __dmul_rn((double)bh_occ(s, i), t.logp[i]));
    return tag;
}

__global__ void k_tags(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states,
                       double* __restrict__ tags, int* __restrict__ iota)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= D) return;
    tags[k] = bh_tag(t, states[k]);
    iota[k] = (int)k;
}

__global__ void k_invert_perm(int64_t D, const int* __restrict__ perm, int* __restrict__ inv)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < D) inv[perm[k]] = (int)k;
}

// Export: column `pos` of the m x D matrix of doubles = occupations of state lex(pos); tags likewise.
__global__ void k_export_basis(const BhTables* __restrict__ gtab, int64_t D, const uint64_t* __restrict__ states,
                               const int* __restrict__ perm /* pos -> lex, or NULL */, double* __restrict__ tags,
                               double* __restrict__ basis)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= D) return;
    const int64_t k = perm ? perm[pos] : pos;
    const uint64_t s = states[k];
    if (tags) tags[pos] = bh_tag(t, s);
    if (basis)
        for (int i = 0; i < t.m; ++i) basis[pos * t.m + i] = (double)bh_occ(s, i);
}

// Rank lookup for arbitrary occupation vectors given as doubles (replaces calculate_tag + search_tag).
__global__ void k_rank_states(const BhTables* __restrict__ gtab, int64_t count, const double* __restrict__ st,
                              const int* __restrict__ inv /* lex -> pos, or NULL */, int* __restrict__ out)
{
    __shared__ BhTables t;
    bh_stage_tables(&t, gtab);
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    int total = 0;
    bool ok = true;
    uint64_t s = 0;
    for (int i = 0; i < t.m; ++i) {
        const double v = st[c * t.m + i];
        const int ni = (int)v;
        if (v < 0 || (double)ni != v || ni > t.n) ok = false;
        total += ni;
        s |= (uint64_t)(ni & 15) << (4 * i);
    }
    if (!ok || total != t.n) {
        out[c] = -1;
        return;
    }
    const int r = bh_rank_of(t, s);
    out[c] = inv ? inv[r] : r;
}

__global__ void k_gather_vec(int64_t D, const int* __restrict__ idx, const double* __restrict__ src,
                             double* __restrict__ dst)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < D) dst[k] = src[idx[k]];
}

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

static const int kPrimes[25] = {2,  3,  5,  7,  11, 13, 17, 19, 23, 29, 31, 37, 41,
                                43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97};

int bh_build_basis(bh_ctx* ctx)
{
    BhTables& t = ctx->h_tab;
    std::memset(&t, 0, sizeof(t));
    const int m = ctx->m, n = ctx->n;
    t.m = m;
    t.n = n;
    for (int q = 0; q < m - 1; ++q)
        for (int R = 0; R <= n + 1; ++R) t.f[q][R] = (R > 0) ? (int)binom64(R - 1 + m - 1 - q, m - 1 - q) : 0;
    for (int q = 0; q < m - 1; ++q)
        for (int R = 0; R <= n; ++R) {
            t.gh[q][R].x = (R >= 1 ? t.f[q][R - 1] : t.f[q][R]) - t.f[q][R];
            t.gh[q][R].y = t.f[q][R + 1] - t.f[q][R];
        }
    for (int a = 0; a < 256; ++a) t.sq[a] = std::sqrt((double)a);
    for (int i = 0; i < m; ++i) t.logp[i] = std::log(kPrimes[i]);
    int nb = 0;
    for (int i = 0; i < m; ++i)
        for (int jj = ctx->nbr_ptr[i]; jj < ctx->nbr_ptr[i + 1]; ++jj) {
            const int src = ctx->nbr_idx[jj];
            // reference fill_hopping: triplets (index,k) and (k,index) for hop src -> i
            t.w[i][src]++;
            t.w[src][i]++;
        }
    for (int src = 0; src < m; ++src)
        for (int dst = 0; dst < m; ++dst)
            if (t.w[dst][src] > 0) t.bond[nb++] = (unsigned short)(dst | (src << 4) | (t.w[dst][src] << 8));
    t.nbonds = nb;
    // recognise the reference's chains (src/neighbours.cpp:21-34): the chain-specialised H.v kernel needs no prefix arrays
    t.chain = 0;
    if (m >= 3) {
        bool open_ok = true, closed_ok = true;
        for (int a = 0; a < m; ++a)
            for (int b2 = 0; b2 < m; ++b2) {
                const bool nn = (a - b2 == 1 || b2 - a == 1);
                const bool wrap = (a == 0 && b2 == m - 1) || (a == m - 1 && b2 == 0);
                if (t.w[a][b2] != (nn ? 2 : 0)) open_ok = false;
                if (t.w[a][b2] != ((nn || wrap) ? 2 : 0)) closed_ok = false;
            }
        t.chain = closed_ok ? 2 : (open_ok ? 1 : 0);
    }
    ctx->max_row = nb + 1;

    BH_CUDA(ctx, cudaMalloc(&ctx->d_tab, sizeof(BhTables)));
    BH_H2D(ctx, ctx->d_tab, &t, sizeof(BhTables));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_states, sizeof(uint64_t) * std::max<int64_t>(ctx->nloc, 1)));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_dU, sizeof(double) * std::max<int64_t>(ctx->nloc, 1)));
    if (ctx->nloc > 0)
        k_unrank<<<nblocks(ctx->nloc, 256), 256, 0, ctx->stream>>>(ctx->d_tab, ctx->row0, ctx->nloc, ctx->d_states, ctx->d_dU);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}

int bh_ensure_orderings(bh_ctx* ctx)
{
    if (ctx->d_perm_tag) return BH_OK;
    const int64_t D = ctx->D;
    double* tags_in = nullptr;
    int* iota = nullptr;
    BH_CUDA(ctx, cudaMalloc(&tags_in, sizeof(double) * D));
    BH_CUDA(ctx, cudaMalloc(&iota, sizeof(int) * D));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_tags, sizeof(double) * D));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_perm_tag, sizeof(int) * D));
    BH_CUDA(ctx, cudaMalloc(&ctx->d_inv_tag, sizeof(int) * D));
    k_tags<<<nblocks(D, 256), 256, 0, ctx->stream>>>(ctx->d_tab, D, ctx->d_states, tags_in, iota);
    BH_LAUNCHED(ctx);
    // tags are >= 0, so the unsigned order of the bit patterns is the numeric order; equal tags cannot
    // occur (distinct states have distinct prime products) and the sort is stable anyway.
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, reinterpret_cast<const uint64_t*>(tags_in),
                                    reinterpret_cast<uint64_t*>(ctx->d_tags), iota, ctx->d_perm_tag, (int)D, 0, 64,
                                    ctx->stream);
    void* tmp = nullptr;
    BH_CUDA(ctx, cudaMalloc(&tmp, tmp_bytes));
    BH_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, reinterpret_cast<const uint64_t*>(tags_in),
                                                 reinterpret_cast<uint64_t*>(ctx->d_tags), iota, ctx->d_perm_tag,
                                                 (int)D, 0, 64, ctx->stream));
    ctx->launches += 8;
    k_invert_perm<<<nblocks(D, 256), 256, 0, ctx->stream>>>(D, ctx->d_perm_tag, ctx->d_inv_tag);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    cudaFree(tags_in);
    cudaFree(iota);
    return BH_OK;
}

// pos -> lex map of an ordering (NULL for LEX) and its inverse
static const int* perm_of(bh_ctx* ctx, int order) { return order == BH_ORDER_TAG_SORTED ? ctx->d_perm_tag : ctx->d_inv_tag; }
static const int* inv_of(bh_ctx* ctx, int order) { return order == BH_ORDER_TAG_SORTED ? ctx->d_inv_tag : ctx->d_perm_tag; }

int bh_permute_vec(bh_ctx* ctx, int order, bool to_order, const double* src, double* dst)
{
    if (order == BH_ORDER_LEX) {
        if (src != dst)
            BH_CUDA(ctx, cudaMemcpyAsync(dst, src, sizeof(double) * ctx->nloc, cudaMemcpyDeviceToDevice, ctx->stream));
        return BH_OK;
    }
    if (ctx->user_matrix) return bh_fail(ctx, BH_ERR_STATE, "a loaded matrix has only its own ordering (BH_ORDER_LEX)");
    BH_TRY(bh_ensure_orderings(ctx));
    // to_order: dst[pos] = src[lex(pos)];  from order: dst[lex] = src[pos(lex)]
    const int* idx = to_order ? perm_of(ctx, order) : inv_of(ctx, order);
    k_gather_vec<<<nblocks(ctx->D, 256), 256, 0, ctx->stream>>>(ctx->D, idx, src, dst);
    BH_LAUNCHED(ctx);
    BH_CUDA(ctx, cudaGetLastError());
    return BH_OK;
}

static int setup_impl(bh_ctx* ctx, int m, int n, const int* nbr_ptr, const int* nbr_idx, bool partition)
{
    if (!ctx) return BH_ERR_ARG;
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    if (m < 1 || m > BH_MAX_SITES || n < 1 || n > BH_MAX_BOSONS)
        return bh_fail(ctx, BH_ERR_UNSUPPORTED, "bh_setup: need 1 <= m <= 16 sites and 1 <= n <= 15 bosons");
    if (!nbr_ptr || (nbr_ptr[m] > 0 && !nbr_idx)) return bh_fail(ctx, BH_ERR_ARG, "bh_setup: neighbour list is NULL");
    int64_t D = 0;
    bh_dimension(m, n, &D);
    if (D >= (int64_t)1 << 31) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "bh_setup: D >= 2^31");
    for (int i = 0; i < m; ++i)
        for (int jj = nbr_ptr[i]; jj < nbr_ptr[i + 1]; ++jj) {
            if (nbr_idx[jj] < 0 || nbr_idx[jj] >= m) return bh_fail(ctx, BH_ERR_ARG, "bh_setup: neighbour out of range");
            if (nbr_idx[jj] == i) return bh_fail(ctx, BH_ERR_ARG, "bh_setup: a site cannot be its own neighbour");
        }
    if (nbr_ptr[m] > 120) return bh_fail(ctx, BH_ERR_UNSUPPORTED, "bh_setup: more than 120 neighbour entries");
    bh_release_system(ctx);
    ctx->m = m;
    ctx->n = n;
    ctx->D = D;
    ctx->nbr_ptr.assign(nbr_ptr, nbr_ptr + m + 1);
    ctx->nbr_idx.assign(nbr_idx, nbr_idx + nbr_ptr[m]);
    if (partition && ctx->world > 1) {
        // row partition for one large eigensolve: equal slices of the LEX rank range (the all-gather needs equal counts)
        const int64_t per = ((D + ctx->world - 1) / ctx->world + 31) / 32 * 32;
        ctx->row0 = std::min<int64_t>(D, per * ctx->rank);
        ctx->nloc = std::min<int64_t>(D, ctx->row0 + per) - ctx->row0;
        ctx->ld = per;
        ctx->partitioned = true;
        BH_TRY(bh_build_basis(ctx));
        BH_CUDA(ctx, cudaMalloc(&ctx->d_xfull, sizeof(double) * per * ctx->world));
        BH_CUDA(ctx, cudaMemsetAsync(ctx->d_xfull, 0, sizeof(double) * per * ctx->world, ctx->stream));
        BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        BH_TRY(bh_dist_plan_halo(ctx));  // chains: which parts of the other slices this rank's hops read (dist.cu)
        return BH_OK;  // matrix-free only: no stored Hamiltonian
    }
    ctx->row0 = 0;
    ctx->nloc = D;
    ctx->ld = (D + 31) / 32 * 32;
    BH_TRY(bh_build_basis(ctx));  // the stored Hamiltonian (K2) is built on first use
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BH_OK;
}

extern "C" int bh_setup(bh_ctx* ctx, int m, int n, const int* nbr_ptr, const int* nbr_idx)
{
    return setup_impl(ctx, m, n, nbr_ptr, nbr_idx, false);
}

extern "C" int bh_setup_partitioned(bh_ctx* ctx, int m, int n, const int* nbr_ptr, const int* nbr_idx)
{
    if (!ctx) return BH_ERR_ARG;
    if (ctx->world < 2) return bh_fail(ctx, BH_ERR_STATE, "bh_setup_partitioned: call bh_dist_init first");
    return setup_impl(ctx, m, n, nbr_ptr, nbr_idx, true);
}

extern "C" int bh_partition(const bh_ctx* ctx, int64_t* row0, int64_t* nrows, int64_t* slice)
{
    if (!ctx || !ctx->D) return BH_ERR_STATE;
    if (row0) *row0 = ctx->row0;
    if (nrows) *nrows = ctx->nloc;
    if (slice) *slice = ctx->ld;
    return BH_OK;
}

int bh_ensure_staging(bh_ctx* ctx)
{
    if (!ctx->D) return bh_fail(ctx, BH_ERR_STATE, "no system: call bh_setup first");
    if (!ctx->d_x) BH_CUDA(ctx, cudaMalloc(&ctx->d_x, sizeof(double) * ctx->ld));
    if (!ctx->d_y) BH_CUDA(ctx, cudaMalloc(&ctx->d_y, sizeof(double) * ctx->ld));
    return BH_OK;
}

extern "C" int bh_basis(bh_ctx* ctx, int order, double* tags, double* basis)
{
    if (!ctx || !ctx->D || ctx->user_matrix || ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_basis: call bh_setup first");
    if (order < 0 || order > 2) return bh_fail(ctx, BH_ERR_ARG, "bh_basis: bad order");
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t D = ctx->D;
    const int* perm = nullptr;
    if (order != BH_ORDER_LEX) {
        BH_TRY(bh_ensure_orderings(ctx));
        perm = perm_of(ctx, order);
    }
    double *d_t = nullptr, *d_b = nullptr;
    if (tags) BH_CUDA(ctx, cudaMalloc(&d_t, sizeof(double) * D));
    if (basis) BH_CUDA(ctx, cudaMalloc(&d_b, sizeof(double) * D * ctx->m));
    k_export_basis<<<nblocks(D, 256), 256, 0, ctx->stream>>>(ctx->d_tab, D, ctx->d_states, perm, d_t, d_b);
    BH_LAUNCHED(ctx);
    if (tags) BH_D2H(ctx, tags, d_t, sizeof(double) * D);
    if (basis)
        BH_D2H(ctx, basis, d_b, sizeof(double) * D * ctx->m);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_dev(d_t);
    free_dev(d_b);
    return BH_OK;
}

extern "C" int bh_rank(bh_ctx* ctx, int order, const double* states, int64_t count, int32_t* ranks)
{
    if (!ctx || !ctx->D || ctx->user_matrix || ctx->partitioned) return bh_fail(ctx, BH_ERR_STATE, "bh_rank: call bh_setup first");
    if (order < 0 || order > 2 || !states || !ranks || count < 0) return bh_fail(ctx, BH_ERR_ARG, "bh_rank: bad argument");
    if (count == 0) return BH_OK;
    BH_CUDA(ctx, cudaSetDevice(ctx->device));
    const int* inv = nullptr;
    if (order != BH_ORDER_LEX) {
        BH_TRY(bh_ensure_orderings(ctx));
        inv = inv_of(ctx, order);
    }
    double* d_s = nullptr;
    int* d_r = nullptr;
    BH_CUDA(ctx, cudaMalloc(&d_s, sizeof(double) * count * ctx->m));
    BH_CUDA(ctx, cudaMalloc(&d_r, sizeof(int) * count));
    BH_H2D(ctx, d_s, states, sizeof(double) * count * ctx->m);
    k_rank_states<<<nblocks(count, 256), 256, 0, ctx->stream>>>(ctx->d_tab, count, d_s, inv, d_r);
    BH_LAUNCHED(ctx);
    BH_D2H(ctx, ranks, d_r, sizeof(int) * count);
    BH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_s);
    cudaFree(d_r);
    return BH_OK;
}

// ---- geometry (host; replaces class Neighbours) ----
extern "C" int bh_neighbours_rect(int lx, int ly, int lz, int closed, int* nbr_ptr, int* nbr_idx)
{
    if (lx < 1 || ly < 1 || lz < 1 || !nbr_ptr) return BH_ERR_ARG;
    // site = (z*ly + y)*lx + x; per axis of extent > 1: "minus" neighbour then "plus" neighbour, wrapping only
    // when closed -- the entry order of Neighbours::{chain,square,cube}_neighbours for their own shapes.
    const int ext[3] = {lx, ly, lz};
    const int stride[3] = {1, lx, lx * ly};
    const int m = lx * ly * lz;
    int pos = 0;
    for (int s = 0; s < m; ++s) {
        nbr_ptr[s] = pos;
        const int c[3] = {s % lx, (s / lx) % ly, s / (lx * ly)};
        for (int a = 0; a < 3; ++a) {
            if (ext[a] < 2) continue;
            for (int dir = -1; dir <= 1; dir += 2) {
                int cc = c[a] + dir;
                if (cc < 0 || cc >= ext[a]) {
                    if (!closed) continue;
                    cc = (cc + ext[a]) % ext[a];
                }
                if (nbr_idx) nbr_idx[pos] = s + (cc - c[a]) * stride[a];
                ++pos;
            }
        }
    }
    nbr_ptr[m] = pos;
    return BH_OK;
}

extern "C" int bh_neighbours_chain(int m, int closed, int* nbr_ptr, int* nbr_idx)
{
    if (m < 1 || !nbr_ptr) return BH_ERR_ARG;
    // src/neighbours.cpp:21-34 appends the periodic partners last: site 0 = {1, m-1}, site m-1 = {m-2, 0}
    int pos = 0;
    for (int i = 0; i < m; ++i) {
        nbr_ptr[i] = pos;
        if (i > 0) { if (nbr_idx) nbr_idx[pos] = i - 1; ++pos; }
        if (i < m - 1) { if (nbr_idx) nbr_idx[pos] = i + 1; ++pos; }
        if (closed && i == 0) { if (nbr_idx) nbr_idx[pos] = m - 1; ++pos; }
        if (closed && i == m - 1) { if (nbr_idx) nbr_idx[pos] = 0; ++pos; }
    }
    nbr_ptr[m] = pos;
    return BH_OK;
}
