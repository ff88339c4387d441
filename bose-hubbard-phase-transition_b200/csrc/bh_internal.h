// bh_internal.h -- private definitions shared by the translation units of libbh_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/bh_b200.h"

#define BH_MAX_SITES 16   // packed state = 16 nibbles in one 64-bit word
#define BH_MAX_BOSONS 15  // one nibble per site
#define BH_MAX_NCV 512    // Krylov basis columns held by the Lanczos workspace

// ---- ranking / lattice tables, one copy in device global memory, staged to shared memory by kernels ----
struct BhTables {
    int m, n;
    // f[q][R] = [R > 0] * C(R - 1 + m - 1 - q, m - 1 - q): number of basis states that precede, on site q,
    // a state with R bosons on the sites after q (descending-lexicographic rank = sum_q f[q][R_q]).
    int f[BH_MAX_SITES][BH_MAX_BOSONS + 3];
    // rank change of a nearest-neighbour hop across bond (q, q+1) when R bosons sit beyond site q:
    // gh[q][R].x = f[q][R-1] - f[q][R] (boson moves q+1 -> q), .y = f[q][R+1] - f[q][R] (boson moves q -> q+1)
    int2 gh[BH_MAX_SITES][BH_MAX_BOSONS + 3];
    // w[dst][src]: how many times the ordered bond appears in the neighbour list, both directions summed
    // (the reference pushes (index,k) and (k,index) for every listed neighbour, src/hamiltonian.cpp:184-185).
    unsigned char w[BH_MAX_SITES][BH_MAX_SITES];
    int chain;   // 1 = open chain, 2 = closed chain (every nearest-neighbour bond listed from both ends, w = 2), else 0
    int nbonds;  // ordered pairs with w > 0
    // the same pairs as a list, ordered by (src, dst): dst | src << 4 | w << 8
    unsigned short bond[BH_MAX_SITES * (BH_MAX_SITES - 1)];
    // sq[a] = sqrt(a) for the amplitudes sqrt((n_dst + 1) * n_src), a <= 16 * 15
    double sq[256];
    double logp[BH_MAX_SITES];  // log(prime_i), host glibc values (src/hamiltonian.cpp:94,144)
};

struct BhProfRec {
    cudaEvent_t a, b;
    int cls;
    double bytes;
};

struct bh_ctx {
    int device = 0;
    // per-kernel event timing (bh_ctx_profile_*): records live on the root context (lockstep children share it)
    bool prof_on = false;
    std::vector<BhProfRec> prof;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[BH_PROF_NCLASSES] = {0}, prof_bytes[BH_PROF_NCLASSES] = {0};
    int64_t prof_n[BH_PROF_NCLASSES] = {0};
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;  // bytes copied across PCIe by this context

    // lockstep batching of grid points (batch.cu): bh_points solves `batch` points together, sharing their H.v launches
    int batch = 1;                       // env BH_BATCH / bh_ctx_set_batch (1 = off, 2..4)
    int batch_plain = 0;                 // 1: the plain H.v of stage 1 / stage 3 go through the lockstep hub too (env BH_BATCH_PLAIN; measured slower: it fragments the filter batches)
    struct bh_batch_hub* hub = nullptr;  // parent side: scheduling state + interleaved buffers
    std::vector<bh_ctx*> children;       // parent side: one workspace-owning context per lockstep solve
    bh_ctx* parent = nullptr;            // child side
    int fiber = -1;                      // child side: index inside the batch

    // system
    bool user_matrix = false;  // true after bh_load_matrix: no Fock basis, only the SELL copy of the given matrix
    int m = 0, n = 0;
    int64_t D = 0;
    int64_t ld = 0;  // padded (local) vector length (multiple of 32 doubles)
    // row partition (one large eigensolve over several GPUs): this context owns LEX ranks [row0, row0 + nloc)
    int world = 1, rank = 0;
    void* nccl_comm = nullptr;
    bool partitioned = false;
    int64_t row0 = 0, nloc = 0;
    double* d_xfull = nullptr;  // all-gathered vector, world * ld doubles (global index = LEX rank)
    // overlapped halo exchange (dist.cu): only the chunks of the other slices that this rank's hops read are exchanged
    // (ncclSend / ncclRecv on a second communicator and stream) while the own-slice hops are computed
    struct HaloRange { int peer; int64_t off, count; };  // off = global element offset
    std::vector<HaloRange> halo_send, halo_recv;
    void* nccl_comm2 = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_x_ready = nullptr, ev_halo_done = nullptr;
    bool halo_ready = false;
    int64_t halo_recv_elems = 0;
    // peer-memory form (dist.cu, default for chains): every vector that can be the input of an H.v (Krylov basis, w, f, the three
    // Chebyshev buffers) lives in ONE arena per rank, exported with CUDA IPC and mapped by every other rank; after a barrier the
    // copy engines pull the ranges of the halo plan straight out of the owners' arenas over NVLink while the SMs sweep the
    // local hops.
    double* d_arena = nullptr;
    int arena_ncv = 0;
    std::vector<void*> peer_arena;   // [world]; own entry = d_arena
    bool peer_ready = false;
    double* d_barrier = nullptr;
    std::vector<cudaStream_t> pull_stream;  // one per peer: the pulls from different peers run on different copy engines
    std::vector<cudaEvent_t> pull_done;
    // the hops of this rank's rows whose source element lives in another rank's slice, stored once as a CSR matrix
    // (pattern and amplitudes are fixed by the basis and the partition; 2J is applied at run time)
    int* d_rem_ptr = nullptr;     // [nloc + 1]
    int* d_rem_col = nullptr;     // global LEX rank of the source
    double* d_rem_amp = nullptr;  // sqrt((n_dst + 1) n_src)
    int64_t rem_nnz = 0;
    std::vector<int> nbr_ptr, nbr_idx;
    BhTables h_tab;
    BhTables* d_tab = nullptr;
    int max_row = 0;  // max entries per row of H (incl. diagonal)
    // SELL-32 copy of the stored H: slices of 32 consecutive rows, column-major inside a slice,
    // padded to the longest row of the slice (padding = zero value pointing at the row's own column)
    int64_t sell_nslices = 0, sell_entries = 0;
    int sell_sigma = 256;           // sorting window (rows), multiple of 32 (env BH_SELL_SIGMA; 32 = plain SELL-32)
    int* d_sell_row = nullptr;      // [nslices * 32] slot -> row (-1 = padding slot)
    int* d_sell_ptr = nullptr;      // [nslices + 1] entry offset of each slice
    int* d_sell_col = nullptr;
    double* d_sell_valJ = nullptr;
    double* d_sell_valH = nullptr;
    int* d_sell_diag = nullptr;     // [D] position of each row's diagonal entry
    bool sell_valid = false, sell_partial_valid = false;
    double sell_cJ = 0, sell_cU = 0, sell_cmu = 0;

    uint64_t* d_states = nullptr;  // packed occupations, LEX order
    double* d_dU = nullptr;        // sum_i n_i (n_i + 1)
    // H pattern = JH u diagonal, LEX order, ascending columns
    int64_t nnzH = 0, nnzJ = 0;
    int* d_rowptr = nullptr;
    int* d_col = nullptr;
    double* d_valJ = nullptr;  // J = 1 values (0 on the diagonal slot)
    int* d_diagpos = nullptr;  // position of the diagonal entry of each row
    double* d_valH = nullptr;  // values of the currently materialised H(cJ,cU,cmu)
    double cur_cJ = 0, cur_cU = 0, cur_cmu = 0;
    bool valH_valid = false;

    // orderings (lazy): perm_tag[i] = LEX index of the i-th smallest tag, inv_tag = inverse
    double* d_tags = nullptr;
    int* d_perm_tag = nullptr;
    int* d_inv_tag = nullptr;

    int cheb_degree = 8;   // Chebyshev filter degree of the accelerated solver (1 = plain Lanczos; env BH_CHEB_DEGREE)
    int cheb_quick = 24;   // quick stage 1 for D >= 50000: one cycle of this many plain steps locates E_0 (0 = always the full stage 1; env BH_CHEB_QUICK)
    int cheb_pre = 3;      // plain restart cycles run first to locate the wanted end of the spectrum (env BH_CHEB_PRE)
    double cheb_margin = 0.05;  // cut >= theta_{nev-1} + margin * (theta_{nev-1} - theta_0)   (env BH_CHEB_MARGIN)
    double cheb_frac = 0.05;    // cut >= theta_0 + frac * (hi - theta_0)                       (env BH_CHEB_FRAC)
    int cheb_stall_iter = 0;    // > 0 while a quick-mode stage 2 runs: give up once this many restarts have left the nev-th Ritz
                                // value of the filtered operator inside the damped band (> -1): the cut sits below E_{nev-1}
    double* d_cheb[3] = {nullptr, nullptr, nullptr};
    int rr_gram = 1;                 // Rayleigh-Ritz of H: W = H V, then one Gram pass V^T W (env BH_RR_GRAM; 0 = ncv transposed products)
    double* d_hv_block = nullptr;    // W, hv_block_cols columns of ld doubles (lazy)
    int hv_block_cols = 0;
    double* d_gram_part = nullptr;   // per-CTA partial Gram matrices
    int compress_tiled = 2;  // restart GEMM (env BH_COMPRESS_TILED): 0 shared-memory rows, 1 4x4 register tiles, 2 4x8 register tiles
    int coop_fused = 1;    // cooperative step: update with block k fused with the dot products of block k+1 (env BH_COOP_FUSED)
    int coop = 1;          // single cooperative launch per Lanczos step when the residual fits in registers (env BH_COOP)
    int reorth_block = 8;  // basis columns per re-orthogonalisation block (env BH_REORTH_BLOCK)
    bool reorth_block_forced = false;
    void* small_ws = nullptr;  // bh_small_ws (small.cu): Krylov workspaces of the many-point small-system solver
    // Lanczos workspace (lazy)
    int ws_ncv = 0;
    double* d_V = nullptr;      // (ws_ncv + 1) columns of ld doubles
    double* d_w = nullptr;      // work vector
    double* d_f = nullptr;      // residual vector
    double* d_scal = nullptr;   // device scalars (see lanczos.cu)
    double* d_part = nullptr;   // per-block partial sums
    unsigned int* d_counter = nullptr;
    double* d_small = nullptr;  // small matrices uploaded from the host (Y, coefficients)
    double* d_spdm_scratch = nullptr;  // SPDM partials + result
    size_t spdm_scratch_bytes = 0;
    double* d_x = nullptr;      // staging vectors for host-pointer entry points
    double* d_y = nullptr;
    double* h_pinned = nullptr;  // pinned host staging
    size_t h_pinned_bytes = 0;
};

// ---- error helpers ----
int bh_fail(bh_ctx* ctx, int code, const std::string& msg);
#define BH_CUDA(ctx, expr)                                                                              \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return bh_fail((ctx), BH_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)
#define BH_TRY(expr)                 \
    do {                             \
        int _rc = (expr);            \
        if (_rc != BH_OK) return _rc; \
    } while (0)
#define BH_LAUNCHED(ctx) ((ctx)->launches++)
// counted host<->device copies on the context's stream
#define BH_H2D(ctx, dst, src, bytes)                                                                      \
    do {                                                                                                  \
        (ctx)->h2d_bytes += (int64_t)(bytes);                                                             \
        BH_CUDA((ctx), cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, (ctx)->stream));    \
    } while (0)
#define BH_D2H(ctx, dst, src, bytes)                                                                      \
    do {                                                                                                  \
        (ctx)->d2h_bytes += (int64_t)(bytes);                                                             \
        BH_CUDA((ctx), cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, (ctx)->stream));    \
    } while (0)

// ---- per-kernel event timing: BhProfScope brackets the launches issued while it is alive ----
void bh_prof_begin(bh_ctx* ctx, int cls, double bytes);
void bh_prof_end(bh_ctx* ctx);
struct BhProfScope {
    bh_ctx* root;
    BhProfScope(bh_ctx* ctx, int cls, double bytes) : root(ctx->parent ? ctx->parent : ctx)
    {
        if (root->prof_on) bh_prof_begin(root, cls, bytes); else root = nullptr;
    }
    ~BhProfScope() { if (root) bh_prof_end(root); }
};

// ---- internal entry points (defined across the .cu files) ----
int bh_release_system(bh_ctx* ctx);
void bh_release_workspace(bh_ctx* ctx);  // Krylov workspace, staging, scratch (everything a lockstep child owns)
// lockstep batching (batch.cu)
bool bh_batch_supported(const bh_ctx* ctx, int kernel);
int bh_points_lockstep(bh_ctx* ctx, int nb, int64_t npoints, const double* cJ, const double* cU, const double* cmu, int nb_eigen,
                       int kernel, double* out3, bh_eigs_info* infos);
int bh_batch_filter(bh_ctx* child, const double* x, double* y, double c, double e, double cJ, double cU, double cmu, int d,
                    bool* handled);
void bh_batch_release(bh_ctx* ctx);
// many grid points of a small system, one CTA per point (small.cu)
bool bh_small_supported(const bh_ctx* ctx, int kernel, int64_t npoints, int nb_eigen);
int bh_points_small(bh_ctx* ctx, int64_t npoints, const double* cJ, const double* cU, const double* cmu, int nb_eigen, double* out3,
                    bh_eigs_info* infos);
void bh_small_release(bh_ctx* ctx);
int bh_build_basis(bh_ctx* ctx);          // K1: states, dU
int bh_build_hamiltonian(bh_ctx* ctx);    // K2: pattern, J values
int bh_ensure_orderings(bh_ctx* ctx);     // tags, radix sort, permutations
int bh_materialise_H(bh_ctx* ctx, double cJ, double cU, double cmu);
int bh_materialise_sell(bh_ctx* ctx, double cJ, double cU, double cmu, int64_t max_entries = -1, int64_t max_rows = -1);  // builds the SELL copy on first use
// y = s1 * (H x) + s2 * x + s3 * z   (the epilogue is fused into the H.v kernels; plain H.v = {1, 0, 0, NULL})
struct BhEpilogue {
    double s1 = 1.0, s2 = 0.0, s3 = 0.0;
    const double* z = nullptr;
};
int bh_launch_hv(bh_ctx* ctx, double cJ, double cU, double cmu, int kernel, const double* x_dev, double* y_dev,
                 const BhEpilogue& ep = BhEpilogue());
// rigorous (Gershgorin) bounds of the spectrum of H(cJ, cU, cmu); model contexts only
int bh_spectrum_bounds(bh_ctx* ctx, double cJ, double cU, double cmu, double* lo, double* hi);
int bh_dist_allreduce_max(bh_ctx* ctx, double* buf_dev, int64_t count);
int bh_ensure_workspace(bh_ctx* ctx, int ncv);
int bh_ensure_staging(bh_ctx* ctx);
// permute between LEX (device) and `order`: dst[pos] = src[lex(pos)] (to_order) or dst[lex(pos)] = src[pos]
int bh_permute_vec(bh_ctx* ctx, int order, bool to_order, const double* src_dev, double* dst_dev);
// eigensolve leaving the Ritz data on the device; see lanczos.cu
struct BhSolve {
    std::vector<double> evals;  // nev ascending
    std::vector<double> Y;      // ncv x nev Ritz coefficient vectors (column-major)
    int ncv = 0, nev = 0;
    bh_eigs_info info{};
};
int bh_lanczos(bh_ctx* ctx, double cJ, double cU, double cmu, int nev, int ncv, double tol, int maxit, int kernel,
               BhSolve* out);
// x_dev = V * Y[:, col]  (Ritz vector in LEX order)
int bh_ritz_vector(bh_ctx* ctx, const BhSolve& s, int col, double* x_dev);
int bh_spdm_dev(bh_ctx* ctx, const double* phi_dev, int ncols, double* rho_host);
// NCCL plumbing (dist.cu; libnccl is dlopen'ed on first use)
int bh_dist_allreduce_sum(bh_ctx* ctx, double* buf_dev, int64_t count);
int bh_dist_allgather(bh_ctx* ctx, const double* send_dev, double* recv_dev, int64_t count_per_rank);
int bh_mark_halo_chunks(bh_ctx* ctx, unsigned char* flags_dev);  // hv.cu
int bh_build_remote_hops(bh_ctx* ctx);                           // hv.cu: d_rem_* (chains, partitioned)
int bh_exclusive_scan(bh_ctx* ctx, int64_t n, const int* d_in, int* d_out, int64_t* total);  // csr.cu: out[0..n]
int bh_dist_plan_halo(bh_ctx* ctx);                              // once per bh_setup_partitioned (chains)
int bh_dist_halo_begin(bh_ctx* ctx, const double* x_local);      // start the exchange of x into d_xfull (communication stream)
int bh_dist_halo_end(bh_ctx* ctx);                               // the context's stream waits for it
void bh_dist_release_halo(bh_ctx* ctx);
bool bh_dist_peer_wanted(const bh_ctx* ctx);                     // partitioned chain context with the peer-memory form enabled
int bh_dist_arena(bh_ctx* ctx, int ncv);                         // (re)allocate + export + open the vector arenas (collective)
void bh_dist_arena_release(bh_ctx* ctx);
int bh_dist_barrier(bh_ctx* ctx);                                // all ranks' prior work on their context streams is complete
int bh_dist_pull_begin(bh_ctx* ctx, int64_t x_off);             // copy-engine pulls of the halo ranges out of the peers' arenas
int bh_dist_pull_end(bh_ctx* ctx);                               // the context's stream waits for them

// host-side small dense symmetric eigen-decomposition (ascending; vectors in columns of v, column-major)
void bh_sym_eig(int n, std::vector<double>& a, std::vector<double>& evals, std::vector<double>& v);
