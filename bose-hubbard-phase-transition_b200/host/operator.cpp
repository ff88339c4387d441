#include "operator.hpp"

#include <mutex>
#include <stdexcept>
#include <vector>

#include "../../include/bh_b200.h"

namespace {
std::mutex g_mtx;
bh_ctx* g_ctx = nullptr;
int g_device = 0;
}  // namespace

void Op::set_device(int d) { std::lock_guard<std::mutex> lk(g_mtx); g_device = d; }

Eigen::VectorXcd Op::IRLM_eigen(Eigen::SparseMatrix<double> O, int nb_eigen, Eigen::MatrixXcd& eigenvectors)
{
    std::lock_guard<std::mutex> lk(g_mtx);
    if (!g_ctx && bh_ctx_create(g_device, &g_ctx) != BH_OK) throw std::runtime_error(bh_last_error(nullptr));
    O.makeCompressed();
    const int64_t D = O.rows();
    if (bh_load_matrix(g_ctx, D, O.outerIndexPtr(), O.innerIndexPtr(), O.valuePtr()) != BH_OK)
        throw std::runtime_error(bh_last_error(g_ctx));
    std::vector<double> evals(nb_eigen), vecs((size_t)D * nb_eigen);
    const int rc = bh_eigs(g_ctx, 0, 0, 0, nb_eigen, 2 * nb_eigen + 1, 1e-10, 1000, BH_HV_USER, BH_ORDER_LEX, evals.data(),
                           vecs.data(), nullptr);
    if (rc == BH_ERR_ARG) throw std::invalid_argument(bh_last_error(g_ctx));
    if (rc != BH_OK) throw std::runtime_error("Eigenvalue computation failed.");
    Eigen::VectorXcd out(nb_eigen);
    for (int i = 0; i < nb_eigen; ++i) out[i] = std::complex<double>(evals[i], 0.0);
    eigenvectors = Eigen::Map<const Eigen::MatrixXd>(vecs.data(), D, nb_eigen).cast<std::complex<double>>();
    return out;
}
