/* bh_b200.h -- C ABI of libbh_b200.so, the B200-native exact-diagonalisation hot path of
 * Bose-Hubbard-Phase-Transition (Fock basis -> Hamiltonian -> lowest eigenpairs -> ground-state
 * observables, swept over a (J, U, mu) grid).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no exceptions.
 * The reference has no FFI layer; each entry point below names the reference interface it
 * replaces (paths relative to the reference repository root).  The C++ shim under
 * bose-hubbard-phase-transition_b200/host/ keeps the reference's own signatures (namespaces BH, Op,
 * Analysis, class Neighbours, the CLI) on top of this ABI; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every function returns BH_OK (0) or a negative BH_ERR_* code; bh_last_error() gives the text;
 *  - all pointers are HOST pointers unless the name ends in _dev (device pointers, used by the
 *    benchmark and by callers that keep vectors resident in HBM);
 *  - there is NO CPU fallback: without a usable CUDA device bh_ctx_create fails;
 *  - a context owns one GPU and one "system" (m sites, n bosons, neighbour list) at a time; calls on
 *    one context are serialised by the caller (the reference's OpenMP threads map to one context
 *    per thread/GPU);
 *  - `order` selects the basis ordering used for every per-state array crossing the boundary:
 *      BH_ORDER_LEX          descending-lexicographic enumeration order (src/hamiltonian.cpp:60-85),
 *                            the order used internally on the device;
 *      BH_ORDER_TAG_SORTED   ascending prime-log tag (what BH::sort_basis intends, src/hamiltonian.cpp:109-123);
 *      BH_ORDER_REF_SCATTER  the order the unmodified BH::fixed_set_basis really returns (its in-place
 *                            permutation applies the inverse).
 *  - matrices are returned in compressed-column form with ascending inner indices, exactly the arrays of
 *    an Eigen::SparseMatrix<double> (outerIndexPtr / innerIndexPtr / valuePtr); H is symmetric so this is
 *    also its CSR form.
 */
#ifndef BH_B200_H
#define BH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bh_ctx bh_ctx;

enum { BH_OK = 0, BH_ERR_ARG = -1, BH_ERR_CUDA = -2, BH_ERR_STATE = -3, BH_ERR_NOCONV = -4, BH_ERR_UNSUPPORTED = -5 };
enum { BH_ORDER_LEX = 0, BH_ORDER_TAG_SORTED = 1, BH_ORDER_REF_SCATTER = 2 };
enum { BH_TERM_J = 0, BH_TERM_U = 1, BH_TERM_MU = 2 };
/* H.v kernels: stored CSR (K3) or matrix-free on-the-fly (K4) */
enum { BH_HV_STORED = 0, BH_HV_MATRIX_FREE = 1, BH_HV_USER = 2 /* matrix loaded with bh_load_matrix */,
       BH_HV_HYBRID = 3 /* chains: first rows through the stored slices, the rest matrix-free, in one launch */ };

/* ---- context ------------------------------------------------------------------------------- */
int bh_ctx_create(int device, bh_ctx** ctx);
int bh_ctx_destroy(bh_ctx* ctx);
/* text of the last error on this context (ctx may be NULL: error of the last failed bh_ctx_create) */
const char* bh_last_error(const bh_ctx* ctx);
/* use an existing CUDA stream (cudaStream_t) for every launch of this context; NULL = own stream */
int bh_ctx_set_stream(bh_ctx* ctx, void* cuda_stream);
/* number of kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t bh_ctx_launch_count(const bh_ctx* ctx);
/* bytes this context has copied host->device and device->host since creation (bench.py's e2e accounting) */
int bh_ctx_transfer_bytes(const bh_ctx* ctx, int64_t* h2d, int64_t* d2h);

/* ---- geometry: replaces class Neighbours (include/neighbours.hpp:12-63, src/neighbours.cpp) ----- */
/* Fill a CSR-style neighbour list: nbr_ptr[m+1], nbr_idx[nbr_ptr[m]].  Call with nbr_idx == NULL to get
 * the sizes only.  Entry order per site follows the reference (left, right[, up, down[, front, back]]). */
int bh_neighbours_chain(int m, int closed, int* nbr_ptr, int* nbr_idx);                    /* src/neighbours.cpp:21-34 */
int bh_neighbours_rect(int lx, int ly, int lz, int closed, int* nbr_ptr, int* nbr_idx);    /* :40-120, any lx*ly*lz */

/* ---- Fock basis: replaces BH::dimension / BH::fixed_set_basis / BH::search_tag -------------- */
int bh_dimension(int m, int n, int64_t* D);                                                /* src/hamiltonian.cpp:39-54 */
/* Build the system on the device: basis, ranking tables, diagonals, hopping matrix (K1, K2). */
int bh_setup(bh_ctx* ctx, int m, int n, const int* nbr_ptr, const int* nbr_idx);
/* tags[D], basis[m*D] column-major (state k = column k), src/hamiltonian.cpp:143-149.  Either may be NULL. */
int bh_basis(bh_ctx* ctx, int order, double* tags, double* basis);
/* ranks[count] = index (in `order`) of each of `count` occupation vectors given as columns of
 * states[m*count] (doubles, like the reference's basis); -1 if it is not a basis state.
 * Replaces calculate_tag + search_tag (src/hamiltonian.cpp:91-97,126-140). */
int bh_rank(bh_ctx* ctx, int order, const double* states, int64_t count, int32_t* ranks);

/* ---- Hamiltonian: replaces BH::fixed_bosons_hamiltonian (src/hamiltonian.cpp:170-256) -------- */
int bh_term_nnz(bh_ctx* ctx, int term, int64_t* nnz);
/* One term scaled by coef, exactly the matrix the reference builds for (J,0,0), (0,U,0) or (0,0,mu). */
int bh_term_csc(bh_ctx* ctx, int term, double coef, int order, int32_t* outer, int32_t* inner, double* val);
/* H = JH*cJ + UH*cU + uH*cmu with the union pattern and explicit zeros of src/analysis.cpp:311;
 * nnz(H) = nnz(JH) + D. */
int bh_hamiltonian_nnz(bh_ctx* ctx, int64_t* nnz);
int bh_hamiltonian_csc(bh_ctx* ctx, double cJ, double cU, double cmu, int order, int32_t* outer, int32_t* inner,
                       double* val);

/* ---- H.v: replaces Spectra::SparseGenMatProd::perform_op (MatOp/SparseGenMatProd.h:81-86) ---- */
/* y = (JH*cJ + UH*cU + uH*cmu) x with host vectors in `order` (copies included: this is the MatOp seam). */
int bh_hv(bh_ctx* ctx, double cJ, double cU, double cmu, int kernel, int order, const double* x, double* y);
/* Page-lock a caller-owned host buffer (cudaHostRegister) so that the copies of bh_hv / bh_eigs / bh_spdm run as direct DMA
 * at PCIe speed instead of through the driver's pageable staging (2 x 10.8 MB at m = n = 12: ~0.45 ms instead of ~1.8 ms per
 * bh_hv).  Optional; the caller unregisters before freeing the buffer. */
int bh_host_register(void* ptr, int64_t bytes);
int bh_host_unregister(void* ptr);
/* Same with device vectors in LEX order, launched on the context's stream, no synchronisation. */
int bh_hv_dev(bh_ctx* ctx, double cJ, double cU, double cmu, int kernel, const double* x_dev, double* y_dev);

/* ---- eigensolver: replaces Op::IRLM_eigen (src/operator.cpp:22-33) --------------------------- */
typedef struct bh_eigs_info {
    int32_t nconv;     /* converged wanted pairs */
    int32_t nmatvec;   /* H.v applications */
    int32_t nrestart;  /* restarts (Spectra's num_iterations) */
    int32_t nreorth;   /* Lanczos steps that ran the full re-orthogonalisation pass */
    double seconds;    /* device+host wall time of the solve */
} bh_eigs_info;
/* nev smallest eigenvalues (ascending) of H(cJ,cU,cmu); thick-restart Lanczos with Spectra's
 * parameters (ncv, tol, maxit; HermEigsBase.h:360-385).  evecs may be NULL, else D*nev column-major in
 * `order`.  kernel selects the H.v implementation.  Returns BH_ERR_NOCONV if fewer than nev converged
 * (the reference throws std::runtime_error("Eigenvalue computation failed.")), BH_ERR_ARG when
 * nev/ncv violate Spectra's constructor checks (nev + 2 <= ncv <= D for the general solver). */
int bh_eigs(bh_ctx* ctx, double cJ, double cU, double cmu, int nev, int ncv, double tol, int maxit, int kernel,
            int order, double* evals, double* evecs, bh_eigs_info* info);

/* ---- generic operator: the literal Op::IRLM_eigen(Eigen::SparseMatrix<double> O, ...) seam -------- */
/* Load any real symmetric sparse matrix (compressed columns = compressed rows, ascending inner indices, the
 * arrays of an Eigen::SparseMatrix<double>) into the context; it replaces the context's system.  The matrix is
 * copied to the device and converted to the SELL-32 layout there.  Afterwards bh_eigs / bh_hv with
 * kernel = BH_HV_USER and order = BH_ORDER_LEX (the matrix's own ordering) operate on it; cJ/cU/cmu are ignored. */
int bh_load_matrix(bh_ctx* ctx, int64_t D, const int32_t* outer, const int32_t* inner, const double* val);

/* ---- observables: replaces Analysis::SPDM/braket/coherence/gap_ratios (src/analysis.cpp:433-594) */
/* rho[m*m] column-major from a state vector phi (host, `order`), divided by ncols (the reference divides
 * by eigenvectors.cols() = nb_eigen, src/analysis.cpp:527). */
int bh_spdm(bh_ctx* ctx, int order, const double* phi, int ncols, double* rho);
int bh_gap_ratios(const double* evals, int nb_eigen, double* ratios /* nb_eigen-2 */);   /* :433-454 */
int bh_condensate_fraction(int m, const double* rho, double* out);                       /* :331-334 */
int bh_coherence(int m, const double* rho, double* out);                                 /* :542-556 */

/* Finite-temperature branch of the sweep body (src/analysis.cpp:321-323, 456-494; dead code in the reference, whose
 * temperature is the constant 0).  Boltzmann weights of the nb_eigen levels with the reference's normalisation (eigenvalues
 * divided by their maximum before the exponential, :481-486); and the dense matrix the reference builds from them,
 * rho = sum_k w_k u_k u_k^T (D x D, column-major, host) from D x nb_eigen eigenvectors (column-major, host; any basis order:
 * the result is in the same order).  O(D^2) memory by construction: D up to a few 10^4. */
int bh_thermal_weights(const double* evals, int nb_eigen, double temperature, double* weights);
int bh_density_matrix(bh_ctx* ctx, int64_t D, int nb_eigen, const double* evals, const double* evecs, double temperature, double* rho);

/* ---- sweep: replaces the body and loop of Analysis::calculate_and_save (src/analysis.cpp:266-387) */
/* One grid point: eigensolve + gap ratio + SPDM + condensate fraction + coherence.
 * out3 = {gap_ratio, condensate_fraction, coherence}; evals[nb_eigen] and rho[m*m] optional. */
int bh_point(bh_ctx* ctx, double cJ, double cU, double cmu, int nb_eigen, int kernel, double* out3, double* evals,
             double* rho, bh_eigs_info* info);
/* A list of grid points on this context (the shard of one GPU): cJ/cU/cmu[npoints] -> out3[3*npoints]. */
int bh_points(bh_ctx* ctx, const double* cJ, const double* cU, const double* cmu, int64_t npoints, int nb_eigen,
              int kernel, double* out3, bh_eigs_info* infos /* may be NULL */);
/* Lockstep batching of bh_points (the independent iterations of the reference's `omp parallel for` over grid points,
 * src/analysis.cpp:302): with batch = 2..4, groups of that many points are solved together and share their H.v
 * launches (closed chains, kernel = BH_HV_MATRIX_FREE; anything else falls back to one point at a time).
 * Results are the same point by point.  Default 1, or the environment variable BH_BATCH. */
int bh_ctx_set_batch(bh_ctx* ctx, int batch);

/* ---- one large eigensolve row-partitioned over several GPUs (BASELINE.json config 5) ---------------------
 * One process (or thread) per GPU.  Rank 0 creates a 128-byte NCCL id (bh_dist_unique_id) and hands it to the
 * others by any means (the Python driver broadcasts it with torch.distributed); every rank calls bh_dist_init on
 * its context, then bh_setup_partitioned: the context owns an equal slice of the LEX rank range, holds only
 * that slice of every vector, applies H matrix-free after an ncclAllGather of the Lanczos vector and
 * ncclAllReduces the recurrence scalars.  bh_eigs / bh_point then work as on one GPU with
 * kernel = BH_HV_MATRIX_FREE and order = BH_ORDER_LEX; eigenvectors come back as the local slice
 * (nrows rows per column, column stride = nrows).  No collective is issued outside these calls. */
int bh_dist_unique_id(void* id128);
int bh_dist_init(bh_ctx* ctx, int world, int rank, const void* id128);
int bh_dist_finalize(bh_ctx* ctx);
int bh_setup_partitioned(bh_ctx* ctx, int m, int n, const int* nbr_ptr, const int* nbr_idx);
int bh_partition(const bh_ctx* ctx, int64_t* row0, int64_t* nrows, int64_t* slice);

/* ---- per-kernel timing (bench.py's roofline entries) ---------------------------------------------------
 * With profiling enabled every launch of the classes below is bracketed by a pair of CUDA events on the context's
 * stream (the launching stream); bh_ctx_profile_read synchronises, adds up launches / milliseconds / algorithmic
 * bytes per class since the last read and clears the records.  Off by default (two event records per launch). */
enum { BH_PROF_HV_FREE = 0 /* matrix-free H.v, one vector */, BH_PROF_HV_BATCH2 = 1, BH_PROF_HV_BATCH4 = 2 /* lockstep H.v */,
       BH_PROF_HV_STORED = 3, BH_PROF_STEP = 4 /* Lanczos step after the H.v: three-term update + re-orthogonalisation */,
       BH_PROF_RESTART = 5 /* V <- V Y */, BH_PROF_GRAM = 6, BH_PROF_SPDM = 7, BH_PROF_SMALL = 8 /* many-point small-system steps */,
       BH_PROF_NCLASSES = 9 };
int bh_ctx_profile_enable(bh_ctx* ctx, int on);
int bh_ctx_profile_read(bh_ctx* ctx, int cls, int64_t* launches, double* total_ms, double* total_bytes);

/* ---- benchmark helpers (device-resident, used by bench.py) --------------------------------------- */
/* Fill x_dev[D] with Spectra's LCG(seed 0) uniform(-0.5,0.5) sequence (Util/SimpleRandom.h:30-64), LEX order. */
int bh_lcg_fill_dev(bh_ctx* ctx, double* x_dev, int64_t count);
/* Algorithmic bytes of one H.v: stored = 12*nnz(H) + 4*(D+1) + 16*D, matrix-free = 16*D. */
int bh_hv_algorithmic_bytes(bh_ctx* ctx, int kernel, int64_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* BH_B200_H */
