// small_dense.cpp -- host-side dense symmetric eigen-decomposition for the projected problems.
//
// The Krylov-projected matrix (at most BH_MAX_NCV x BH_MAX_NCV: tridiagonal, or arrowhead + tridiagonal after
// a thick restart) and the m x m single-particle density matrix are solved on the host, as the north star
// prescribes ("the small tridiagonal eigensolve stays on the host").  Replaces Spectra's TridiagEigen
// (LinAlg/TridiagEigen.h) and the Eigen::EigenSolver call of src/analysis.cpp:331.
//
// Method: Householder reduction to tridiagonal form with accumulated transformations, then the implicit
// symmetric QR iteration with Wilkinson shifts (Golub & Van Loan, Matrix Computations, alg. 8.3.1-8.3.3).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "bh_internal.h"

void bh_sym_eig(int n, std::vector<double>& a, std::vector<double>& evals, std::vector<double>& vec)
{
    // a: n x n column-major symmetric (destroyed); evals ascending; vec: eigenvectors in columns
    auto A = [&](int i, int j) -> double& { return a[i + (size_t)j * n]; };
    std::vector<double> q((size_t)n * n, 0.0);
    auto Q = [&](int i, int j) -> double& { return q[i + (size_t)j * n]; };
    for (int i = 0; i < n; ++i) Q(i, i) = 1.0;
    std::vector<double> v(n), p(n), t(n);

    // ---- Householder tridiagonalisation, column by column ----
    for (int k = 0; k + 2 < n; ++k) {
        double tail = 0;
        for (int i = k + 2; i < n; ++i) tail += A(i, k) * A(i, k);
        if (tail == 0) continue;
        const double x0 = A(k + 1, k);
        const double norm = std::sqrt(x0 * x0 + tail);
        const double alpha = (x0 > 0) ? -norm : norm;
        double vn = 0;
        for (int i = k + 1; i < n; ++i) {
            v[i] = A(i, k);
            if (i == k + 1) v[i] -= alpha;
            vn += v[i] * v[i];
        }
        vn = std::sqrt(vn);
        for (int i = k + 1; i < n; ++i) v[i] /= vn;
        // p = A_sub v ; K = v.p ; w = p - K v ; A_sub -= 2 (v w^T + w v^T)
        double K = 0;
        for (int i = k + 1; i < n; ++i) {
            double s = 0;
            const double* col = &A(0, i);  // A is symmetric: row i = column i, contiguous
            for (int j = k + 1; j < n; ++j) s += col[j] * v[j];
            p[i] = s;
            K += s * v[i];
        }
        for (int i = k + 1; i < n; ++i) p[i] -= K * v[i];
        for (int j = k + 1; j < n; ++j)
            for (int i = k + 1; i < n; ++i) A(i, j) -= 2.0 * (v[i] * p[j] + p[i] * v[j]);
        A(k + 1, k) = A(k, k + 1) = alpha;
        for (int i = k + 2; i < n; ++i) A(i, k) = A(k, i) = 0.0;
        // Q <- Q (I - 2 v v^T), column by column (contiguous): t = 2 Q v, then Q(:, j) -= t v_j
        std::fill(t.begin(), t.end(), 0.0);
        for (int j = k + 1; j < n; ++j) {
            const double vj = 2.0 * v[j];
            const double* col = &Q(0, j);
            for (int i = 0; i < n; ++i) t[i] += col[i] * vj;
        }
        for (int j = k + 1; j < n; ++j) {
            const double vj = v[j];
            double* col = &Q(0, j);
            for (int i = 0; i < n; ++i) col[i] -= t[i] * vj;
        }
    }
    std::vector<double> d(n), e(std::max(n - 1, 0));
    for (int i = 0; i < n; ++i) d[i] = A(i, i);
    for (int i = 0; i + 1 < n; ++i) e[i] = A(i + 1, i);

    // ---- implicit symmetric QR with Wilkinson shift ----
    const double eps = 2.220446049250313e-16;
    int hi = n - 1;
    int guard = 0;
    while (hi > 0 && guard < 100 * n) {
        for (int i = 0; i < hi; ++i)
            if (std::fabs(e[i]) <= eps * (std::fabs(d[i]) + std::fabs(d[i + 1]))) e[i] = 0.0;
        if (e[hi - 1] == 0.0) {
            --hi;
            continue;
        }
        int lo = hi - 1;
        while (lo > 0 && e[lo - 1] != 0.0) --lo;
        ++guard;
        const double dd = (d[hi - 1] - d[hi]) / 2.0;
        const double eh = e[hi - 1];
        const double den = dd + (dd >= 0 ? 1.0 : -1.0) * std::sqrt(dd * dd + eh * eh);
        const double mu = d[hi] - eh * eh / den;
        double x = d[lo] - mu, z = e[lo];
        for (int k = lo; k < hi; ++k) {
            const double r = std::sqrt(x * x + z * z);  // entries are O(|T|): no overflow guard needed (std::hypot costs more than the rotation)
            const double c = (r == 0) ? 1.0 : x / r;
            const double s = (r == 0) ? 0.0 : -z / r;
            if (k > lo) e[k - 1] = r;
            const double d1 = d[k], d2 = d[k + 1], ek = e[k];
            d[k] = c * c * d1 - 2 * c * s * ek + s * s * d2;
            d[k + 1] = s * s * d1 + 2 * c * s * ek + c * c * d2;
            e[k] = c * s * (d1 - d2) + (c * c - s * s) * ek;
            if (k < hi - 1) {
                x = e[k];
                z = -s * e[k + 1];
                e[k + 1] = c * e[k + 1];
            }
            double* __restrict__ q0 = &Q(0, k);
            double* __restrict__ q1 = &Q(0, k + 1);
            for (int i = 0; i < n; ++i) {
                const double qk = q0[i], qk1 = q1[i];
                q0[i] = c * qk - s * qk1;
                q1[i] = s * qk + c * qk1;
            }
        }
    }
    std::vector<int> ind(n);
    std::iota(ind.begin(), ind.end(), 0);
    std::stable_sort(ind.begin(), ind.end(), [&](int x1, int x2) { return d[x1] < d[x2]; });
    evals.resize(n);
    vec.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        evals[j] = d[ind[j]];
        for (int i = 0; i < n; ++i) vec[i + (size_t)j * n] = Q(i, ind[j]);
    }
}
