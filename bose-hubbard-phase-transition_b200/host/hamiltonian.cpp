#include "hamiltonian.hpp"

#include <algorithm>
#include <cmath>
#include <iostream>
#include <limits>
#include <mutex>
#include <numeric>
#include <stdexcept>

namespace {

std::mutex g_mtx;
bh_ctx* g_ctx = nullptr;
int g_device = 0;
int g_order = BH_ORDER_TAG_SORTED;
int g_m = 0, g_n = 0;
std::vector<int> g_ptr, g_idx;

void check(int rc, bh_ctx* ctx)
{
    if (rc == BH_OK) return;
    const std::string msg = bh_last_error(ctx);
    if (rc == BH_ERR_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

// (re)build the device system when (m, n, neighbour list) changes
bh_ctx* system_for(int m, int n, const std::vector<std::vector<int>>* nei)
{
    if (!g_ctx) check(bh_ctx_create(g_device, &g_ctx), nullptr);
    std::vector<int> ptr(m + 1, 0), idx;
    if (nei) {
        for (int i = 0; i < m; ++i) {
            ptr[i + 1] = ptr[i] + static_cast<int>((*nei)[i].size());
            idx.insert(idx.end(), (*nei)[i].begin(), (*nei)[i].end());
        }
    } else if (g_m == m && g_n == n) {
        return g_ctx;  // basis only: any lattice will do
    } else {
        idx.resize(2 * m + 2);
        bh_neighbours_chain(m, 1, ptr.data(), idx.data());
        idx.resize(ptr[m]);
    }
    if (g_m == m && g_n == n && ptr == g_ptr && idx == g_idx) return g_ctx;
    if (idx.empty()) idx.push_back(0);
    check(bh_setup(g_ctx, m, n, ptr.data(), idx.data()), g_ctx);
    g_m = m; g_n = n; g_ptr = ptr; g_idx = idx;
    if (g_idx.size() == 1 && ptr[m] == 0) g_idx.clear();
    return g_ctx;
}

// which of the three orderings does this basis have?  (-1: none)
int detect_order(bh_ctx* ctx, const Eigen::MatrixXd& basis)
{
    const int64_t D = basis.cols();
    std::vector<int32_t> r(D);
    for (int order : {BH_ORDER_TAG_SORTED, BH_ORDER_REF_SCATTER, BH_ORDER_LEX}) {
        check(bh_rank(ctx, order, basis.data(), D, r.data()), ctx);
        bool same = true;
        for (int64_t i = 0; i < D && same; ++i) same = (r[i] == i);
        if (same) return order;
    }
    return -1;
}

}  // namespace

void BH::set_basis_order(int o) { std::lock_guard<std::mutex> lk(g_mtx); g_order = o; }
void BH::set_device(int d) { std::lock_guard<std::mutex> lk(g_mtx); g_device = d; }

double BH::calculate_tag(const Eigen::MatrixXd& basis, const std::vector<int>& primes, int k)
{
    double tag = 0;
    for (int i = 0; i < basis.rows(); i++) tag += basis.coeff(i, k) * std::log(primes[i]);
    return tag;
}

Eigen::VectorXd BH::calculate_tags(const Eigen::MatrixXd& basis, const std::vector<int>& primes)
{
    Eigen::VectorXd tags(basis.cols());
    for (int i = 0; i < basis.cols(); i++) tags[i] = calculate_tag(basis, primes, i);
    return tags;
}

int BH::search_tag(const Eigen::VectorXd& tags, double x)
{
    int a = 0, b = static_cast<int>(tags.size()) - 1, mid = (a + b) / 2;
    while (std::fabs(tags[mid] - x) > 1e-12 && a <= b) {
        if (tags[mid] < x) a = mid + 1; else b = mid - 1;
        mid = (a + b) / 2;
    }
    return mid;
}

std::pair<Eigen::VectorXd, Eigen::MatrixXd> BH::fixed_set_basis(int m, int n)
{
    std::lock_guard<std::mutex> lk(g_mtx);
    bh_ctx* ctx = system_for(m, n, nullptr);
    int64_t D = 0;
    bh_dimension(m, n, &D);
    Eigen::VectorXd tags(D);
    Eigen::MatrixXd basis(m, D);
    check(bh_basis(ctx, g_order, tags.data(), basis.data()), ctx);
    return std::make_pair(tags, basis);
}

std::pair<Eigen::VectorXd, Eigen::MatrixXd> BH::max_set_basis(int m, int n)
{
    // src/hamiltonian.cpp:152-166: concatenate the fixed-N bases for N = 1..n, then order by tag
    std::vector<double> tags;
    std::vector<Eigen::MatrixXd> blocks;
    int64_t total = 0;
    for (int bosons = 1; bosons <= n; ++bosons) {
        auto tb = fixed_set_basis(m, bosons);
        tags.insert(tags.end(), tb.first.data(), tb.first.data() + tb.first.size());
        total += tb.second.cols();
        blocks.push_back(std::move(tb.second));
    }
    Eigen::MatrixXd all(m, total);
    int64_t off = 0;
    for (auto& b : blocks) { all.middleCols(off, b.cols()) = b; off += b.cols(); }
    std::vector<int64_t> ind(total);
    std::iota(ind.begin(), ind.end(), 0);
    std::sort(ind.begin(), ind.end(), [&](int64_t a, int64_t b) { return tags[a] < tags[b]; });
    Eigen::VectorXd t2(total);
    Eigen::MatrixXd b2(m, total);
    for (int64_t i = 0; i < total; ++i) { t2[i] = tags[ind[i]]; b2.col(i) = all.col(ind[i]); }
    return std::make_pair(t2, b2);
}

Eigen::SparseMatrix<double> BH::fixed_bosons_hamiltonian(const std::vector<std::vector<int>>& neighbours,
                                                         const Eigen::MatrixXd& basis, const Eigen::VectorXd& tags, int m,
                                                         int n, double J, double U, double mu)
{
    (void)tags;
    std::lock_guard<std::mutex> lk(g_mtx);
    int64_t D = 0;
    bh_dimension(m, n, &D);
    Eigen::SparseMatrix<double> H(D, D);
    const double eps = std::numeric_limits<double>::epsilon();
    int term;
    double coef;
    if (std::abs(J - 0.0) > eps) { term = BH_TERM_J; coef = J; }
    else if (std::abs(U - 0.0) > eps) { term = BH_TERM_U; coef = U; }
    else if (std::abs(mu - 0.0) > eps) { term = BH_TERM_MU; coef = mu; }
    else {
        std::cerr << "Error: At least one of the parameters J, U, mu must be different from zero." << std::endl;
        return H;
    }
    bh_ctx* ctx = system_for(m, n, &neighbours);
    if (basis.rows() != m || basis.cols() != D) throw std::invalid_argument("basis must be m x D");
    const int order = detect_order(ctx, basis);
    if (order < 0) throw std::invalid_argument("basis is not in one of the orders returned by BH::fixed_set_basis");
    int64_t nnz = 0;
    check(bh_term_nnz(ctx, term, &nnz), ctx);
    H.resizeNonZeros(nnz);
    check(bh_term_csc(ctx, term, coef, order, H.outerIndexPtr(), H.innerIndexPtr(), H.valuePtr()), ctx);
    return H;
}

Eigen::SparseMatrix<double> BH::max_bosons_hamiltonian(const std::vector<std::vector<int>>& neighbours, int m, int n_min,
                                                       int n_max, double J, double U, double mu)
{
    // src/hamiltonian.cpp:260-288: block diagonal over the boson numbers
    if (n_min < 0) n_min = 0;
    if (n_max < n_min) n_max = n_min;
    std::vector<Eigen::SparseMatrix<double>> blocks;
    int64_t total = 0;
    const double eps = std::numeric_limits<double>::epsilon();
    for (int bosons = n_min; bosons <= n_max; ++bosons) {
        if (bosons == 0) {
            // the empty sector (src/hamiltonian.cpp:268-274 starts at n_min = 0): one state, no hop; the U / mu branches of
            // fixed_bosons_hamiltonian (:248-253) store the explicit diagonal entry U * 0 or -mu * 0.  Built on the host
            // (the device basis starts at one boson).
            Eigen::SparseMatrix<double> h0(1, 1);
            if (std::abs(J) > eps) {
            } else if (std::abs(U) > eps) {
                std::vector<Eigen::Triplet<double>> t{Eigen::Triplet<double>(0, 0, U * 0.0)};
                h0.setFromTriplets(t.begin(), t.end());
            } else if (std::abs(mu) > eps) {
                std::vector<Eigen::Triplet<double>> t{Eigen::Triplet<double>(0, 0, -mu * 0.0)};
                h0.setFromTriplets(t.begin(), t.end());
            }
            blocks.push_back(h0);
            total += 1;
            continue;
        }
        auto tb = fixed_set_basis(m, bosons);
        blocks.push_back(fixed_bosons_hamiltonian(neighbours, tb.second, tb.first, m, bosons, J, U, mu));
        total += blocks.back().rows();
    }
    std::vector<Eigen::Triplet<double>> trip;
    int64_t off = 0;
    for (const auto& h : blocks) {
        for (int k = 0; k < h.outerSize(); ++k)
            for (Eigen::SparseMatrix<double>::InnerIterator it(h, k); it; ++it)
                trip.emplace_back(it.row() + off, it.col() + off, it.value());
        off += h.rows();
    }
    Eigen::SparseMatrix<double> all(total, total);
    all.setFromTriplets(trip.begin(), trip.end());
    return all;
}

void BH::h_MF(double psi, int p, double mu, double J, int q, Eigen::MatrixXd& h)
{
    // occupation basis |0> .. |2p>: diagonal -mu k + k (k - 1) / 2 + q J psi^2, hopping to the mean field -q J psi sqrt(k + 1)
    const int dim = 2 * p + 1;
    const double c = q * J * psi;
    for (int k = 0; k < dim; ++k) h(k, k) = -mu * k + 0.5 * k * (k - 1) + c * psi;
    for (int k = 0; k + 1 < dim; ++k) {
        const double v = -c * std::sqrt(static_cast<double>(k + 1));
        h(k + 1, k) = v;
        h(k, k + 1) = v;
    }
}
