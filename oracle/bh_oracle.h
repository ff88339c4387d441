/* oracle/bh_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU restatement of the reference's exact-diagonalisation
 * path (oracle/bh_oracle.cpp).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (libbh_b200.so) never links or calls anything declared here.
 *
 * "patched" = semantics of the reference with the six patches P1..P6 of
 * SURVEY.md section 8c (UB / race / non-functional code fixed, every
 * deterministic quirk kept).  Citations are relative to /root/reference.
 */
#ifndef BH_ORACLE_H
#define BH_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { BHO_ORDER_LEX = 0, BHO_ORDER_TAG_SORTED = 1, BHO_ORDER_REF_SCATTER = 2 };

/* src/hamiltonian.cpp:39-54  (32-bit int arithmetic of the reference kept) */
int bho_binomial(int n, int k);
int bho_dimension(int m, int n);

/* src/hamiltonian.cpp:60-123,143-149.  tags[D], basis[m*D] column-major (state k = column k).
 * order LEX = init_lexicographic order (no sort), TAG_SORTED = patched sort_basis (gather),
 * REF_SCATTER = unpatched sort_basis (literal cycle walk). */
int bho_basis(int m, int n, int order, double* tags, double* basis);

/* src/hamiltonian.cpp:126-140 with the tolerance as a parameter (1e-3 unpatched, 1e-12 patched) */
int bho_search_tag(const double* tags, int D, double x, double tol);

/* src/hamiltonian.cpp:170-191 (patched P1,P3) + Eigen setFromTriplets semantics
 * (include/Eigen/src/SparseCore/SparseMatrix.h:1035-1063).  Neighbour list in CSR form
 * (nbr_ptr[m+1], nbr_idx).  Two-call: with inner == NULL only the nnz is returned.
 * Returns nnz (>= 0) or -1. */
long bho_hopping_csc(int m, int D, const int* nbr_ptr, const int* nbr_idx, const double* tags,
                     const double* basis, double J, int* outer, int* inner, double* val);

/* src/hamiltonian.cpp:194-232: the two diagonal terms for U = 1 and mu = 1:
 * dU[k] = sum_i n_i (n_i + 1),  dN[k] = -(sum_i n_i)  */
void bho_diagonals(int m, int D, const double* basis, double* dU, double* dN);

/* src/analysis.cpp:311 (-f J mode): H = JH*cJ + UH*cU + uH*cu, union pattern with explicit
 * zeros kept; JH given as CSC for J = 1.  nnz(H) = nnz(JH) + D (JH has no diagonal). */
long bho_hsum_csc(int D, const int* jouter, const int* jinner, const double* jval, const double* dU,
                  const double* dN, double cJ, double cU, double cu, int* outer, int* inner, double* val);

/* include/Eigen/src/SparseCore/SparseDenseProduct.h:86-107 : y = A x, CSC scatter */
void bho_spmv_csc(int D, const int* outer, const int* inner, const double* val, const double* x, double* y);

/* external/spectra/include/Spectra/Util/SimpleRandom.h:30-64, seed 0: uniform(-0.5, 0.5) */
void bho_lcg_vector(long n, double* out);

/* Restatement of Spectra's symmetric implicitly-restarted Lanczos solver
 * (HermEigsBase.h:102-217,360-385; LinAlg/Lanczos.h:59-184; LinAlg/Arnoldi.h:132-180,305-324),
 * selection = smallest algebraic, results ascending.  The reference calls the general
 * (Arnoldi) solver on this symmetric H (src/operator.cpp:22-33); both return the same 20
 * values to ~1e-14 relative and this is re-checked against
 * oracle/_ref in tests/.  evecs may be NULL, else D*nev column-major.
 * Returns the number of converged pairs (nev on success), -1 on bad arguments. */
int bho_eigs_sym(int D, const int* outer, const int* inner, const double* val, int nev, int ncv, double tol,
                 int maxit, double* evals, double* evecs, int* nmatvec, int* nrestart);

/* src/operator.cpp:93-101 (Op::exact_eigen): dense symmetric eigenvalues (cyclic Jacobi), ascending.
 * a is n*n column-major and is destroyed; vecs may be NULL. */
void bho_dense_sym_eig(int n, double* a, double* evals, double* vecs);

/* src/analysis.cpp:433-454; out has nb_eigen-2 entries */
void bho_gap_ratios(const double* evals, int nb_eigen, double* out);

/* src/analysis.cpp:497-528,562-594 (patched P1): rho[m*m] column-major, divided by ncols (20 in the sweep) */
void bho_spdm(int m, int D, const double* tags, const double* basis, const double* phi0, int ncols, double* rho);

/* src/analysis.cpp:331-334 and :542-556 */
double bho_condensate_fraction(int m, const double* rho);
double bho_coherence(int m, const double* rho);

/* One grid point = body of the sweep loop, src/analysis.cpp:311-337 (patched P5,P6):
 * out5 = {p1?, ...} is NOT filled with the parameters; out3 = gap_ratio, condensate_fraction, coherence.
 * evals (nb_eigen) and rho (m*m) are optional. */
int bho_point(int m, int D, const double* tags, const double* basis, const int* jouter, const int* jinner,
              const double* jval, const double* dU, const double* dN, double cJ, double cU, double cu, int nb_eigen,
              double* out3, double* evals, double* rho, int* nmatvec);

#ifdef __cplusplus
}
#endif
#endif
