// oracle/bh_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (plain C++17, no Eigen, no Spectra) of the reference's
// exact-diagonalisation path with the patched semantics P1..P6 of SURVEY.md
// section 8c.  It is the checker of the CUDA product, never the product: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load liboracle.so.
//
// Pinning: tests/test_oracle_vs_ref.py compares every function below with the
// real (patched) reference compiled into oracle/_ref/ where /root/reference
// exists, and tests/test_oracle_golden.py with the committed fixtures under
// tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py).
// The reference's own test-suite pins nothing on this path.
//
// Each function cites the reference file:line it follows (relative to /root/reference).
#include "bh_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace {

const int kPrimes[25] = {2,  3,  5,  7,  11, 13, 17, 19, 23, 29, 31, 37, 41,
                         43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97};  // src/hamiltonian.cpp:144

// src/hamiltonian.cpp:91-97 -- tag = sum_i n_i * log(p_i), sequential in site order,
// log() evaluated per term as the reference does.
double tag_of(const double* state, int m)
{
    double tag = 0;
    for (int i = 0; i < m; i++) tag += state[i] * std::log(kPrimes[i]);
    return tag;
}

// src/hamiltonian.cpp:60-72
bool next_lex(double* s, int m, int n)
{
    for (int k = m - 2; k > -1; k--) {
        if (s[k] != 0) {
            s[k] -= 1;
            int sum = 0;
            for (int i = 0; i <= k; i++) sum += (int)s[i];
            s[k + 1] = n - sum;
            for (int i = k + 2; i < m; i++) s[i] = 0;
            return true;
        }
    }
    return false;
}

// ---- small dense helpers (column-major n x n) ----

// cyclic Jacobi; on exit a holds the eigenvalues on its diagonal, v the eigenvectors (columns)
void jacobi(int n, double* a, double* v)
{
    for (int i = 0; i < n * n; i++) v[i] = 0;
    for (int i = 0; i < n; i++) v[i + i * n] = 1;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0, diag = 0;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) (i == j ? diag : off) += a[i + j * n] * a[i + j * n];
        if (off <= 1e-34 * (diag + off) || off == 0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = a[p + q * n];
                if (apq == 0) continue;
                double app = a[p + p * n], aqq = a[q + q * n];
                double theta = (aqq - app) / (2 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {  // columns p,q of a
                    double akp = a[k + p * n], akq = a[k + q * n];
                    a[k + p * n] = c * akp - s * akq;
                    a[k + q * n] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {  // rows p,q of a
                    double apk = a[p + k * n], aqk = a[q + k * n];
                    a[p + k * n] = c * apk - s * aqk;
                    a[q + k * n] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = v[k + p * n], vkq = v[k + q * n];
                    v[k + p * n] = c * vkp - s * vkq;
                    v[k + q * n] = s * vkp + c * vkq;
                }
            }
    }
}

double dot(const double* a, const double* b, long n)
{
    double s = 0;
    for (long i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
double nrm2(const double* a, long n) { return std::sqrt(dot(a, a, n)); }

struct Csc {
    int D;
    const int* outer;
    const int* inner;
    const double* val;
};

// ---- Spectra symmetric IRL restatement ----
struct Irl {
    Csc A;
    int n, nev, ncv;
    std::vector<double> V;  // n x ncv column-major (Arnoldi.h m_fac_V)
    std::vector<double> H;  // ncv x ncv column-major (m_fac_H)
    std::vector<double> f;  // residual (m_fac_f)
    double beta = 0;
    int k = 0;  // current factorisation size (m_k)
    int nmatop = 0;
    std::vector<double> ritz_val, ritz_est, ritz_vec;  // ncv, ncv, ncv x nev
    std::vector<char> ritz_conv;
    long rng_state = 1;
    const double eps = std::numeric_limits<double>::epsilon();
    const double near0 = std::numeric_limits<double>::min() * 10;

    void matvec(const double* x, double* y)
    {
        bho_spmv_csc(A.D, A.outer, A.inner, A.val, x, y);
        nmatop++;
    }
    double& h(int i, int j) { return H[i + (size_t)j * ncv]; }
    double* col(int j) { return &V[(size_t)j * n]; }

    // LinAlg/Arnoldi.h:132-180
    void init(const double* v0)
    {
        V.assign((size_t)n * ncv, 0);
        H.assign((size_t)ncv * ncv, 0);
        f.assign(n, 0);
        double* v = col(0);
        matvec(v0, v);
        double vn = nrm2(v, n);
        for (int i = 0; i < n; i++) v[i] /= vn;
        std::vector<double> w(n);
        matvec(v, w.data());
        h(0, 0) = dot(v, w.data(), n);
        double mx = 0;
        for (int i = 0; i < n; i++) {
            f[i] = w[i] - v[i] * h(0, 0);
            mx = std::max(mx, std::fabs(f[i]));
        }
        if (mx < eps * std::fabs(h(0, 0))) {
            std::fill(f.begin(), f.end(), 0.0);
            beta = 0;
        } else
            beta = nrm2(f.data(), n);
        k = 1;
    }

    // LinAlg/Arnoldi.h:64-113 (expand_basis): new residual orthogonal to the first i columns
    void expand_basis(int i)
    {
        const double thresh = eps * std::sqrt((double)n);
        std::vector<double> Vf(i);
        for (int iter = 0; iter < 5; iter++) {
            // Spectra seeds a fresh SimpleRandom(seed + 123*iter); any vector outside span(V) serves
            long seed = (2 * i + 123 * iter) & 2147483647L;
            if (!seed) seed = 1;
            for (int r = 0; r < n; r++) {
                seed = (seed * 16807L) % 2147483647L;
                f[r] = (double)seed / 2147483647.0 - 0.5;
            }
            for (int j = 0; j < i; j++) Vf[j] = dot(col(j), f.data(), n);
            for (int j = 0; j < i; j++)
                for (int r = 0; r < n; r++) f[r] -= col(j)[r] * Vf[j];
            beta = nrm2(f.data(), n);
            if (beta < thresh) continue;
            double err = 0;
            int count = 0;
            for (int j = 0; j < i; j++) {
                Vf[j] = dot(col(j), f.data(), n);
                err = std::max(err, std::fabs(Vf[j]));
            }
            while (count < 3 && err > eps * beta) {
                for (int j = 0; j < i; j++)
                    for (int r = 0; r < n; r++) f[r] -= col(j)[r] * Vf[j];
                beta = nrm2(f.data(), n);
                err = 0;
                for (int j = 0; j < i; j++) {
                    Vf[j] = dot(col(j), f.data(), n);
                    err = std::max(err, std::fabs(Vf[j]));
                }
                count++;
            }
            if (err < eps * beta) return;
        }
    }

    // LinAlg/Lanczos.h:59-184
    void factorize_from(int from_k, int to_m)
    {
        if (to_m <= from_k) return;
        const double beta_thresh = eps * std::sqrt((double)n);
        const double eps_sqrt = std::sqrt(eps);
        std::vector<double> Vf(to_m), w(n);
        for (int j = from_k; j < ncv; j++)
            for (int i = 0; i < ncv; i++) h(i, j) = 0;
        for (int i = from_k; i < ncv; i++)
            for (int j = 0; j < from_k; j++) h(i, j) = 0;
        for (int i = from_k; i <= to_m - 1; i++) {
            bool restart = (beta < near0);
            double* v = col(i);
            if (!restart) {
                for (int r = 0; r < n; r++) v[r] = f[r] / beta;
                if (beta < eps_sqrt) {
                    double Viv = dot(col(i - 1), v, n);
                    restart = (std::fabs(Viv) > eps_sqrt);
                }
            }
            if (restart) {
                expand_basis(i);
                for (int r = 0; r < n; r++) v[r] = f[r] / beta;
            }
            h(i, i - 1) = restart ? 0.0 : beta;
            h(i - 1, i) = h(i, i - 1);
            matvec(v, w.data());
            if (!restart) {
                const double b = h(i, i - 1);
                const double* vp = col(i - 1);
                for (int r = 0; r < n; r++) w[r] -= b * vp[r];
            }
            h(i, i) = dot(v, w.data(), n);
            for (int r = 0; r < n; r++) f[r] = w[r] - h(i, i) * v[r];
            beta = nrm2(f.data(), n);
            const int i1 = i + 1;
            double ortho_err = 0;
            for (int j = 0; j < i1; j++) {
                Vf[j] = dot(col(j), f.data(), n);
                ortho_err = std::max(ortho_err, std::fabs(Vf[j]));
            }
            int count = 0;
            while (count < 5 && ortho_err > eps * beta) {
                if (beta < beta_thresh) {
                    std::fill(f.begin(), f.end(), 0.0);
                    beta = 0;
                    break;
                }
                for (int j = 0; j < i1; j++) {
                    const double c = Vf[j];
                    const double* vj = col(j);
                    for (int r = 0; r < n; r++) f[r] -= vj[r] * c;
                }
                h(i - 1, i) += Vf[i - 1];
                h(i, i - 1) = h(i - 1, i);
                h(i, i) += Vf[i];
                beta = nrm2(f.data(), n);
                ortho_err = 0;
                for (int j = 0; j < i1; j++) {
                    Vf[j] = dot(col(j), f.data(), n);
                    ortho_err = std::max(ortho_err, std::fabs(Vf[j]));
                }
                count++;
            }
        }
        k = to_m;
    }

    // HermEigsBase.h:199-217, selection SmallestAlge: ascending Ritz values first
    void retrieve_ritzpair()
    {
        std::vector<double> a(H), vecs((size_t)ncv * ncv);
        jacobi(ncv, a.data(), vecs.data());
        std::vector<int> ind(ncv);
        std::iota(ind.begin(), ind.end(), 0);
        std::sort(ind.begin(), ind.end(), [&](int x, int y) { return a[x + (size_t)x * ncv] < a[y + (size_t)y * ncv]; });
        for (int i = 0; i < ncv; i++) {
            ritz_val[i] = a[ind[i] + (size_t)ind[i] * ncv];
            ritz_est[i] = vecs[(ncv - 1) + (size_t)ind[i] * ncv];
        }
        for (int i = 0; i < nev; i++)
            for (int r = 0; r < ncv; r++) ritz_vec[r + (size_t)i * ncv] = vecs[r + (size_t)ind[i] * ncv];
    }

    // HermEigsBase.h:152-169
    int num_converged(double tol)
    {
        const double eps23 = std::pow(eps, 2.0 / 3);
        int c = 0;
        for (int i = 0; i < nev; i++) {
            double thresh = tol * std::max(eps23, std::fabs(ritz_val[i]));
            double resid = std::fabs(ritz_est[i]) * beta;
            ritz_conv[i] = resid < thresh;
            c += ritz_conv[i];
        }
        return c;
    }

    // HermEigsBase.h:172-196
    int nev_adjusted(int nconv)
    {
        int nev_new = nev;
        for (int i = nev; i < ncv; i++)
            if (std::fabs(ritz_est[i]) < near0) nev_new++;
        nev_new += std::min(nconv, (ncv - nev_new) / 2);
        if (nev_new == 1 && ncv >= 6)
            nev_new = ncv / 2;
        else if (nev_new == 1 && ncv > 2)
            nev_new = 2;
        if (nev_new > ncv - 1) nev_new = ncv - 1;
        return nev_new;
    }

    // HermEigsBase.h:102-148 with LinAlg/UpperHessenbergQR.h (TridiagQR) done as dense Givens QR
    void restart(int knew)
    {
        if (knew >= ncv) return;
        const int nshift = ncv - knew;
        std::vector<double> shifts(ritz_val.begin() + knew, ritz_val.end());
        std::sort(shifts.begin(), shifts.end(), [](double a, double b) { return std::fabs(a) > std::fabs(b); });
        std::vector<double> Q((size_t)ncv * ncv, 0.0);
        for (int i = 0; i < ncv; i++) Q[i + (size_t)i * ncv] = 1;
        std::vector<double> cs(ncv), sn(ncv);
        for (int s = 0; s < nshift; s++) {
            const double mu = shifts[s];
            // R = G_{n-2}..G_0 (H - mu I)
            std::vector<double> R(H);
            for (int i = 0; i < ncv; i++) R[i + (size_t)i * ncv] -= mu;
            for (int i = 0; i < ncv - 1; i++) {
                double a = R[i + (size_t)i * ncv], b = R[(i + 1) + (size_t)i * ncv];
                double r = std::hypot(a, b);
                double c = 1, sgn = 0;
                if (r > 0) { c = a / r; sgn = b / r; }
                cs[i] = c; sn[i] = sgn;
                for (int j = 0; j < ncv; j++) {
                    double x = R[i + (size_t)j * ncv], y = R[(i + 1) + (size_t)j * ncv];
                    R[i + (size_t)j * ncv] = c * x + sgn * y;
                    R[(i + 1) + (size_t)j * ncv] = -sgn * x + c * y;
                }
            }
            // H <- R Q + mu I,  Q_total <- Q_total Q   (Q = G_0^T ... G_{n-2}^T)
            for (int i = 0; i < ncv - 1; i++) {
                double c = cs[i], sg = sn[i];
                for (int r = 0; r < ncv; r++) {
                    double x = R[r + (size_t)i * ncv], y = R[r + (size_t)(i + 1) * ncv];
                    R[r + (size_t)i * ncv] = c * x + sg * y;
                    R[r + (size_t)(i + 1) * ncv] = -sg * x + c * y;
                    double qx = Q[r + (size_t)i * ncv], qy = Q[r + (size_t)(i + 1) * ncv];
                    Q[r + (size_t)i * ncv] = c * qx + sg * qy;
                    Q[r + (size_t)(i + 1) * ncv] = -sg * qx + c * qy;
                }
            }
            for (int i = 0; i < ncv; i++) R[i + (size_t)i * ncv] += mu;
            // keep it exactly symmetric tridiagonal as TridiagQR::matrix_QtHQ does
            std::fill(H.begin(), H.end(), 0.0);
            for (int i = 0; i < ncv; i++) {
                h(i, i) = R[i + (size_t)i * ncv];
                if (i + 1 < ncv) {
                    double o = R[(i + 1) + (size_t)i * ncv];
                    h(i + 1, i) = o;
                    h(i, i + 1) = o;
                }
            }
            k--;
        }
        // LinAlg/Arnoldi.h:305-324 compress_V
        std::vector<double> Vs((size_t)n * (k + 1), 0.0);
        for (int j = 0; j <= k; j++) {
            double* out = &Vs[(size_t)j * n];
            for (int c = 0; c < ncv; c++) {
                double q = Q[c + (size_t)j * ncv];
                if (q == 0) continue;
                const double* vc = col(c);
                for (int r = 0; r < n; r++) out[r] += vc[r] * q;
            }
        }
        std::memcpy(V.data(), Vs.data(), sizeof(double) * (size_t)n * (k + 1));
        const double q = Q[(ncv - 1) + (size_t)(k - 1) * ncv];
        const double hk = h(k, k - 1);
        const double* vk = col(k);
        for (int r = 0; r < n; r++) f[r] = f[r] * q + vk[r] * hk;
        beta = nrm2(f.data(), n);
        factorize_from(k, ncv);
        retrieve_ritzpair();
    }
};

}  // namespace

extern "C" {

int bho_binomial(int n, int k)
{
    if (k == 0 || k == n) return 1;
    if (k > n / 2) return bho_binomial(n, n - k);
    return n * bho_binomial(n - 1, k - 1) / k;
}

int bho_dimension(int m, int n) { return bho_binomial(m + n - 1, n); }

int bho_basis(int m, int n, int order, double* tags, double* basis)
{
    const int D = bho_dimension(m, n);
    // src/hamiltonian.cpp:75-85
    std::vector<double> state(m, 0.0);
    state[0] = n;
    int colc = 0;
    do {
        std::memcpy(basis + (size_t)colc * m, state.data(), sizeof(double) * m);
        colc++;
    } while (next_lex(state.data(), m, n));
    if (colc != D) return -1;
    // src/hamiltonian.cpp:100-106
    for (int i = 0; i < D; i++) tags[i] = tag_of(basis + (size_t)i * m, m);
    if (order == BHO_ORDER_LEX) return 0;
    // src/hamiltonian.cpp:109-123
    std::vector<int> indices(D);
    std::iota(indices.begin(), indices.end(), 0);
    std::sort(indices.begin(), indices.end(), [&](int a, int b) { return tags[a] < tags[b]; });
    if (order == BHO_ORDER_REF_SCATTER) {
        // the literal (unpatched) in-place cycle walk of :115-122
        std::vector<double> tmp(m);
        for (int i = 0; i < D; ++i) {
            while (indices[i] != i) {
                int j = indices[i];
                std::swap(tags[i], tags[j]);
                std::memcpy(tmp.data(), basis + (size_t)i * m, sizeof(double) * m);
                std::memcpy(basis + (size_t)i * m, basis + (size_t)j * m, sizeof(double) * m);
                std::memcpy(basis + (size_t)j * m, tmp.data(), sizeof(double) * m);
                std::swap(indices[i], indices[j]);
            }
        }
        return 0;
    }
    // P2: gather
    std::vector<double> t2(D), b2((size_t)D * m);
    for (int i = 0; i < D; i++) {
        t2[i] = tags[indices[i]];
        std::memcpy(&b2[(size_t)i * m], basis + (size_t)indices[i] * m, sizeof(double) * m);
    }
    std::memcpy(tags, t2.data(), sizeof(double) * D);
    std::memcpy(basis, b2.data(), sizeof(double) * (size_t)D * m);
    return 0;
}

int bho_search_tag(const double* tags, int D, double x, double tol)
{
    int a = 0, b = D - 1, mid = (a + b) / 2;
    while (std::fabs(tags[mid] - x) > tol && a <= b) {
        if (tags[mid] < x)
            a = mid + 1;
        else
            b = mid - 1;
        mid = (a + b) / 2;
    }
    return mid;
}

long bho_hopping_csc(int m, int D, const int* nbr_ptr, const int* nbr_idx, const double* tags, const double* basis,
                     double J, int* outer, int* inner, double* val)
{
    struct Trip { int r, c; double v; };
    std::vector<Trip> trips;
    std::vector<double> state(m);
    for (int k = 0; k < D; k++) {
        const double* bk = basis + (size_t)k * m;
        for (int i = 0; i < m; i++) {
            for (int jj = nbr_ptr[i]; jj < nbr_ptr[i + 1]; jj++) {
                const int src = nbr_idx[jj];  // P3
                std::memcpy(state.data(), bk, sizeof(double) * m);
                if (bk[i] >= 0 && bk[src] >= 1) {
                    state[i] += 1;
                    state[src] -= 1;
                    double x = tag_of(state.data(), m);  // P1
                    int index = bho_search_tag(tags, D, x, 1e-12);  // P4
                    double value = std::sqrt((bk[i] + 1) * bk[src]);
                    trips.push_back({index, k, -J * value});
                    trips.push_back({k, index, -J * value});
                }
            }
        }
    }
    // setFromTriplets (SparseMatrix.h:1035-1063): bucket by row in insertion order, sum duplicates
    // in insertion order (first occurrence keeps the slot), then transpose -> CSC with sorted rows.
    std::vector<long> rstart(D + 1, 0);
    for (auto& t : trips) rstart[t.r + 1]++;
    for (int i = 0; i < D; i++) rstart[i + 1] += rstart[i];
    std::vector<int> rc(trips.size());
    std::vector<double> rv(trips.size());
    {
        std::vector<long> pos(rstart.begin(), rstart.end() - 1);
        for (auto& t : trips) { rc[pos[t.r]] = t.c; rv[pos[t.r]] = t.v; pos[t.r]++; }
    }
    std::vector<int> wi(D, -1);
    std::vector<long> rend(D);
    long count = 0;
    std::vector<long> nstart(D + 1, 0);
    for (int r = 0; r < D; r++) {
        long start = count;
        nstart[r] = start;
        for (long p = rstart[r]; p < rstart[r + 1]; p++) {
            int c = rc[p];
            if (wi[c] >= start) {
                rv[wi[c]] += rv[p];
            } else {
                rv[count] = rv[p];
                rc[count] = c;
                wi[c] = (int)count;
                count++;
            }
        }
    }
    nstart[D] = count;
    const long nnz = count;
    if (!inner) return nnz;
    // transpose (row-bucketed -> column-major with ascending row indices)
    std::vector<long> cstart(D + 1, 0);
    for (long p = 0; p < nnz; p++) cstart[rc[p] + 1]++;
    for (int i = 0; i < D; i++) cstart[i + 1] += cstart[i];
    for (int i = 0; i <= D; i++) outer[i] = (int)cstart[i];
    std::vector<long> pos(cstart.begin(), cstart.end() - 1);
    for (int r = 0; r < D; r++)
        for (long p = nstart[r]; p < nstart[r + 1]; p++) {
            long q = pos[rc[p]]++;
            inner[q] = r;
            val[q] = rv[p];
        }
    return nnz;
}

void bho_diagonals(int m, int D, const double* basis, double* dU, double* dN)
{
    for (int k = 0; k < D; k++) {
        double vU = 0, vN = 0;
        for (int i = 0; i < m; i++) {
            double ni = basis[(size_t)k * m + i];
            vU += (ni + 1) * ni;  // src/hamiltonian.cpp:203-207 with U = 1
            vN += ni;             // src/hamiltonian.cpp:224-228
        }
        dU[k] = 1.0 * vU;
        dN[k] = -1.0 * vN;  // -mu * value with mu = 1
    }
}

long bho_hsum_csc(int D, const int* jouter, const int* jinner, const double* jval, const double* dU, const double* dN,
                  double cJ, double cU, double cu, int* outer, int* inner, double* val)
{
    // Eigen's sparse sum iterates the union of the two patterns; a missing operand counts as 0
    // (include/Eigen/src/SparseCore/SparseCwiseBinaryOp.h): (JH*cJ + UH*cU) + uH*cu
    long q = 0;
    for (int c = 0; c < D; c++) {
        if (outer) outer[c] = (int)q;
        bool placed = false;
        const double dv = (0.0 + dU[c] * cU) + dN[c] * cu;
        for (int p = jouter[c]; p < jouter[c + 1]; p++) {
            int r = jinner[p];
            if (!placed && r > c) {
                if (inner) { inner[q] = c; val[q] = dv; }
                q++;
                placed = true;
            }
            if (r == c) {
                if (inner) { inner[q] = c; val[q] = ((jval[p] * cJ) + dU[c] * cU) + dN[c] * cu; }
                q++;
                placed = true;
                continue;
            }
            if (inner) { inner[q] = r; val[q] = ((jval[p] * cJ) + 0.0) + 0.0; }
            q++;
        }
        if (!placed) {
            if (inner) { inner[q] = c; val[q] = dv; }
            q++;
        }
    }
    if (outer) outer[D] = (int)q;
    return q;
}

void bho_spmv_csc(int D, const int* outer, const int* inner, const double* val, const double* x, double* y)
{
    for (int i = 0; i < D; i++) y[i] = 0;
    for (int j = 0; j < D; j++) {
        const double xj = x[j];
        for (int p = outer[j]; p < outer[j + 1]; p++) y[inner[p]] += val[p] * xj;
    }
}

void bho_lcg_vector(long n, double* out)
{
    // SimpleRandom.h:30-64 : Lehmer generator a = 16807, modulus 2^31-1, seed 0 -> 1
    long seed = 1;
    for (long i = 0; i < n; i++) {
        seed = (seed * 16807L) % 2147483647L;
        out[i] = (double)seed / (double)2147483647L - 0.5;
    }
}

int bho_eigs_sym(int D, const int* outer, const int* inner, const double* val, int nev, int ncv, double tol, int maxit,
                 double* evals, double* evecs, int* nmatvec, int* nrestart)
{
    if (ncv > D) ncv = D;
    if (nev < 1 || nev > D - 1 || ncv <= nev) return -1;  // HermEigsBase.h:269-273
    Irl s;
    s.A = {D, outer, inner, val};
    s.n = D; s.nev = nev; s.ncv = ncv;
    s.ritz_val.assign(ncv, 0); s.ritz_est.assign(ncv, 0); s.ritz_vec.assign((size_t)ncv * nev, 0);
    s.ritz_conv.assign(nev, 0);
    std::vector<double> v0(D);
    bho_lcg_vector(D, v0.data());  // HermEigsBase.h:331-336
    s.init(v0.data());
    // HermEigsBase.h:360-385
    s.factorize_from(1, ncv);
    s.retrieve_ritzpair();
    int i, nconv = 0;
    for (i = 0; i < maxit; i++) {
        nconv = s.num_converged(tol);
        if (nconv >= nev) break;
        s.restart(s.nev_adjusted(nconv));
    }
    // values are already ascending (SmallestAlge selection); final sorting SmallestAlge = P6
    for (int j = 0; j < nev; j++) evals[j] = s.ritz_val[j];
    if (evecs) {
        for (int j = 0; j < nev; j++) {
            double* out = evecs + (size_t)j * D;
            for (int r = 0; r < D; r++) out[r] = 0;
            for (int c = 0; c < ncv; c++) {
                double y = s.ritz_vec[c + (size_t)j * ncv];
                const double* vc = s.col(c);
                for (int r = 0; r < D; r++) out[r] += vc[r] * y;
            }
        }
    }
    if (nmatvec) *nmatvec = s.nmatop;
    if (nrestart) *nrestart = i + 1;
    return std::min(nev, nconv);
}

void bho_dense_sym_eig(int n, double* a, double* evals, double* vecs)
{
    std::vector<double> v((size_t)n * n);
    jacobi(n, a, v.data());
    std::vector<int> ind(n);
    std::iota(ind.begin(), ind.end(), 0);
    std::sort(ind.begin(), ind.end(), [&](int x, int y) { return a[x + (size_t)x * n] < a[y + (size_t)y * n]; });
    for (int i = 0; i < n; i++) {
        evals[i] = a[ind[i] + (size_t)ind[i] * n];
        if (vecs) std::memcpy(vecs + (size_t)i * n, &v[(size_t)ind[i] * n], sizeof(double) * n);
    }
}

void bho_gap_ratios(const double* evals, int nb_eigen, double* out)
{
    std::vector<double> s(evals, evals + nb_eigen);
    std::sort(s.begin(), s.end());
    for (int i = 1; i < nb_eigen - 1; ++i) {
        double lo = std::min(s[i + 1] - s[i], s[i] - s[i - 1]);
        double hi = std::max(s[i + 1] - s[i], s[i] - s[i - 1]);
        out[i - 1] = (hi != 0) ? (lo / hi) : 0;
    }
}

void bho_spdm(int m, int D, const double* tags, const double* basis, const double* phi0, int ncols, double* rho)
{
    const double eps = std::numeric_limits<double>::epsilon();
    std::vector<double> state(m);
    for (int i = 0; i < m; i++) {
        for (int j = i; j < m; j++) {
            // src/analysis.cpp:562-594 (braket)
            double acc = 0;
            for (int k = 0; k < D; k++) {
                if (std::fabs(phi0[k]) > eps) {
                    std::memcpy(state.data(), basis + (size_t)k * m, sizeof(double) * m);
                    if (state[i] >= 0 && state[j] >= 1) {
                        state[i] += 1;
                        state[j] -= 1;
                        double x = tag_of(state.data(), m);
                        int index = bho_search_tag(tags, D, x, 1e-12);
                        if (std::fabs(phi0[index]) > eps) acc += phi0[k] * phi0[index] * std::sqrt(state[i] * state[j]);
                    }
                }
            }
            rho[i + (size_t)j * m] = acc;
        }
        for (int j = 0; j < i; j++) rho[i + (size_t)j * m] = rho[j + (size_t)i * m];
    }
    for (int i = 0; i < m * m; i++) rho[i] /= ncols;  // src/analysis.cpp:527
}

double bho_condensate_fraction(int m, const double* rho)
{
    std::vector<double> a(rho, rho + (size_t)m * m), ev(m);
    double tr = 0;
    for (int i = 0; i < m; i++) tr += rho[i + (size_t)i * m];
    bho_dense_sym_eig(m, a.data(), ev.data(), nullptr);
    double best = ev[0];
    for (int i = 1; i < m; i++)
        if (std::fabs(best) < std::fabs(ev[i])) best = ev[i];
    return std::fabs(best / tr);
}

double bho_coherence(int m, const double* rho)
{
    double sum_all = 0, sum_diag = 0;
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) {
            sum_all += rho[i + (size_t)j * m] * rho[j + (size_t)i * m];
            if (i == j) sum_diag += rho[i + (size_t)j * m] * rho[j + (size_t)i * m];
        }
    return (sum_all - sum_diag) / sum_all;
}

int bho_point(int m, int D, const double* tags, const double* basis, const int* jouter, const int* jinner,
              const double* jval, const double* dU, const double* dN, double cJ, double cU, double cu, int nb_eigen,
              double* out3, double* evals, double* rho, int* nmatvec)
{
    long nnz = bho_hsum_csc(D, jouter, jinner, jval, dU, dN, cJ, cU, cu, nullptr, nullptr, nullptr);
    std::vector<int> outer(D + 1), inner(nnz);
    std::vector<double> val(nnz);
    bho_hsum_csc(D, jouter, jinner, jval, dU, dN, cJ, cU, cu, outer.data(), inner.data(), val.data());
    std::vector<double> ev(nb_eigen), vec((size_t)D * nb_eigen);
    int nconv = bho_eigs_sym(D, outer.data(), inner.data(), val.data(), nb_eigen, 2 * nb_eigen + 1, 1e-10, 1000, ev.data(),
                             vec.data(), nmatvec, nullptr);
    if (nconv < nb_eigen) return -1;  // src/operator.cpp:27-29
    std::vector<double> gr(nb_eigen - 2);
    bho_gap_ratios(ev.data(), nb_eigen, gr.data());
    double g = 0;
    for (double x : gr) g += x;
    out3[0] = gr.empty() ? 0.0 : g / gr.size();
    std::vector<double> r((size_t)m * m);
    bho_spdm(m, D, tags, basis, vec.data(), nb_eigen, r.data());
    out3[1] = bho_condensate_fraction(m, r.data());
    out3[2] = bho_coherence(m, r.data());
    if (evals) std::memcpy(evals, ev.data(), sizeof(double) * nb_eigen);
    if (rho) std::memcpy(rho, r.data(), sizeof(double) * m * m);
    return 0;
}

}  // extern "C"
