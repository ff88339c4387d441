// oracle/spectra_matop_test.cpp -- TEST INFRASTRUCTURE ONLY (built into oracle/_ref/, needs /root/reference headers).
//
// The innermost seam of the reference: Spectra's MatOp concept.  The reference's solver call
// (src/operator.cpp:22-33: Spectra::GenEigsSolver<SparseGenMatProd<double>>(op, nev, 2 nev + 1), compute(SmallestReal))
// is run here UNCHANGED, from Spectra's own headers, on a MatOp whose perform_op is the B200 H.v (bh_hv, host pointers
// in and out) instead of Eigen's sparse product.  Prints the eigenvalues and the residuals of Spectra's eigenvectors
// under the GPU operator as one JSON line.
//
//   spectra_matop_test m n cJ cU cmu nev kernel(0 stored | 1 matrix-free)
#include <Eigen/Dense>
#include <Spectra/GenEigsSolver.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/bh_b200.h"

struct BhMatProd {  // the wrapper INTEGRATION.md section 4 shows
    using Scalar = double;
    bh_ctx* ctx;
    double cJ, cU, cmu;
    int kernel;
    Eigen::Index n;
    mutable long calls = 0;
    Eigen::Index rows() const { return n; }
    Eigen::Index cols() const { return n; }
    void perform_op(const double* x, double* y) const
    {
        ++calls;
        if (bh_hv(ctx, cJ, cU, cmu, kernel, BH_ORDER_TAG_SORTED, x, y) != BH_OK) {
            fprintf(stderr, "bh_hv: %s\n", bh_last_error(ctx));
            exit(3);
        }
    }
};

int main(int argc, char** argv)
{
    if (argc < 8) return 2;
    const int m = atoi(argv[1]), n = atoi(argv[2]);
    const double cJ = atof(argv[3]), cU = atof(argv[4]), cmu = atof(argv[5]);
    const int nev = atoi(argv[6]), kernel = atoi(argv[7]);
    bh_ctx* ctx = nullptr;
    if (bh_ctx_create(0, &ctx) != BH_OK) { fprintf(stderr, "%s\n", bh_last_error(nullptr)); return 3; }
    std::vector<int> ptr(m + 1), idx(2 * m + 2);
    bh_neighbours_chain(m, 1, ptr.data(), idx.data());
    if (bh_setup(ctx, m, n, ptr.data(), idx.data()) != BH_OK) { fprintf(stderr, "%s\n", bh_last_error(ctx)); return 3; }
    int64_t D = 0;
    bh_dimension(m, n, &D);
    BhMatProd op{ctx, cJ, cU, cmu, kernel, (Eigen::Index)D};
    const auto t0 = std::chrono::steady_clock::now();
    Spectra::GenEigsSolver<BhMatProd> eigs(op, nev, 2 * nev + 1);
    eigs.init();
    const int nconv = eigs.compute(Spectra::SortRule::SmallestReal, 1000, 1e-10, Spectra::SortRule::SmallestReal);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (eigs.info() != Spectra::CompInfo::Successful) { printf("{\"ok\": false, \"nconv\": %d}\n", nconv); return 1; }
    Eigen::VectorXcd ev = eigs.eigenvalues();
    Eigen::MatrixXcd vecs = eigs.eigenvectors();
    double worst = 0.0;
    Eigen::VectorXd y(D);
    for (int k = 0; k < nev; ++k) {
        Eigen::VectorXd u = vecs.col(k).real();
        op.perform_op(u.data(), y.data());
        worst = std::max(worst, (y - ev[k].real() * u).cwiseAbs().maxCoeff());
    }
    // the library's own solver on the same operator
    std::vector<double> mine(nev);
    bh_eigs_info info{};
    const int rc = bh_eigs(ctx, cJ, cU, cmu, nev, 2 * nev + 1, 1e-10, 1000, kernel, BH_ORDER_TAG_SORTED, mine.data(), nullptr, &info);
    printf("{\"ok\": true, \"D\": %ld, \"nconv\": %d, \"matop_calls\": %ld, \"spectra_iterations\": %ld, \"seconds\": %.6f, "
           "\"ms_per_perform_op\": %.4f, \"max_residual\": %.3e, \"bh_eigs_rc\": %d, \"evals\": [",
           (long)D, nconv, op.calls, (long)eigs.num_iterations(), secs, 1e3 * secs / (double)op.calls, worst, rc);
    for (int k = 0; k < nev; ++k) printf("%s%.17g", k ? ", " : "", ev[k].real());
    printf("], \"bh_evals\": [");
    for (int k = 0; k < nev; ++k) printf("%s%.17g", k ? ", " : "", mine[k]);
    printf("]}\n");
    bh_ctx_destroy(ctx);
    return 0;
}
