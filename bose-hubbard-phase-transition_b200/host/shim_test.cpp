// shim_test.cpp -- test driver for the Eigen-typed drop-in layer (BH::, Op::, Neighbours).  It accepts the
// same commands and writes the same dump files as oracle/ref_harness.cpp, so tests/test_shim.py drives the
// reference harness and this binary through one code path and compares the dumps.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "hamiltonian.hpp"
#include "neighbours.hpp"
#include "operator.hpp"

static void dump(const std::string& out, const char* name, const double* p, size_t n)
{
    FILE* fp = fopen((out + "." + name + ".f64").c_str(), "wb");
    fwrite(p, sizeof(double), n, fp);
    fclose(fp);
}
static void dump(const std::string& out, const char* name, const int* p, size_t n)
{
    FILE* fp = fopen((out + "." + name + ".i32").c_str(), "wb");
    fwrite(p, sizeof(int), n, fp);
    fclose(fp);
}
static std::vector<std::vector<int>> lattice(const std::string& spec, int m)
{
    Neighbours nb(m);
    int lx = 0, ly = 0;
    if (spec == "chain") nb.chain_neighbours();
    else if (spec == "openchain") nb.chain_neighbours(false);
    else if (sscanf(spec.c_str(), "rect:%d:%d", &lx, &ly) == 2) nb.rect_neighbours(lx, ly);
    else { fprintf(stderr, "bad lattice %s\n", spec.c_str()); exit(2); }
    return nb.getNeighbours();
}
static void dump_csc(const std::string& out, Eigen::SparseMatrix<double>& H)
{
    H.makeCompressed();
    dump(out, "outer", H.outerIndexPtr(), H.outerSize() + 1);
    dump(out, "inner", H.innerIndexPtr(), H.nonZeros());
    dump(out, "val", H.valuePtr(), H.nonZeros());
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    try {
        if (cmd == "basis") {  // basis m n out [scatter]
            if (argc > 5 && std::string(argv[5]) == "scatter") BH::set_basis_order(BH_ORDER_REF_SCATTER);
            auto tb = BH::fixed_set_basis(atoi(argv[2]), atoi(argv[3]));
            dump(argv[4], "tags", tb.first.data(), tb.first.size());
            dump(argv[4], "basis", tb.second.data(), tb.second.size());
            printf("{\"D\": %ld}\n", (long)tb.first.size());
        } else if (cmd == "csc") {  // csc m n term lattice out
            const int m = atoi(argv[2]), n = atoi(argv[3]);
            const std::string term = argv[4];
            auto nei = lattice(argv[5], m);
            auto tb = BH::fixed_set_basis(m, n);
            auto H = BH::fixed_bosons_hamiltonian(nei, tb.second, tb.first, m, n, term == "J" ? 1 : 0, term == "U" ? 1 : 0,
                                                  term == "u" ? 1 : 0);
            dump_csc(argv[6], H);
            printf("{\"D\": %ld, \"nnz\": %ld}\n", (long)H.rows(), (long)H.nonZeros());
        } else if (cmd == "maxbasis") {  // maxbasis m n out
            auto tb = BH::max_set_basis(atoi(argv[2]), atoi(argv[3]));
            dump(argv[4], "tags", tb.first.data(), tb.first.size());
            dump(argv[4], "basis", tb.second.data(), tb.second.size());
            printf("{\"D\": %ld}\n", (long)tb.first.size());
        } else if (cmd == "maxham") {  // maxham m nmin nmax J U u lattice out
            const int m = atoi(argv[2]);
            auto nei = lattice(argv[8], m);
            auto H = BH::max_bosons_hamiltonian(nei, m, atoi(argv[3]), atoi(argv[4]), atof(argv[5]), atof(argv[6]), atof(argv[7]));
            dump_csc(argv[9], H);
            printf("{\"D\": %ld, \"nnz\": %ld}\n", (long)H.rows(), (long)H.nonZeros());
        } else if (cmd == "eigs") {  // eigs m n cJ cU cu nev lattice out : the calls of src/analysis.cpp:231-236,311,314
            const int m = atoi(argv[2]), n = atoi(argv[3]);
            const double cJ = atof(argv[4]), cU = atof(argv[5]), cu = atof(argv[6]);
            const int nev = atoi(argv[7]);
            auto nei = lattice(argv[8], m);
            auto tb = BH::fixed_set_basis(m, n);
            Eigen::SparseMatrix<double> JH = BH::fixed_bosons_hamiltonian(nei, tb.second, tb.first, m, n, 1, 0, 0);
            Eigen::SparseMatrix<double> UH = BH::fixed_bosons_hamiltonian(nei, tb.second, tb.first, m, n, 0, 1, 0);
            Eigen::SparseMatrix<double> uH = BH::fixed_bosons_hamiltonian(nei, tb.second, tb.first, m, n, 0, 0, 1);
            JH = JH * cJ;
            Eigen::SparseMatrix<double> H = JH + UH * cU + uH * cu;
            Eigen::MatrixXcd vecs;
            const auto t0 = std::chrono::steady_clock::now();
            Eigen::VectorXcd ev = Op::IRLM_eigen(H, nev, vecs);
            const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            Eigen::VectorXd e = ev.real();
            dump(argv[9], "evals", e.data(), e.size());
            // residual of every returned pair, checked here with Eigen's own product
            double worst = 0;
            for (int k = 0; k < nev; ++k) {
                Eigen::VectorXd u = vecs.col(k).real();
                worst = std::max(worst, (H * u - e[k] * u).cwiseAbs().maxCoeff());
            }
            dump_csc(std::string(argv[9]) + ".H", H);
            printf("{\"D\": %ld, \"seconds\": %.6f, \"max_residual\": %.3e}\n", (long)H.rows(), secs, worst);
        } else {
            fprintf(stderr, "unknown command\n");
            return 2;
        }
    } catch (const std::invalid_argument& e) {
        printf("{\"exception\": \"invalid_argument\", \"what\": \"%s\"}\n", e.what());
        return 3;
    } catch (const std::runtime_error& e) {
        printf("{\"exception\": \"runtime_error\", \"what\": \"%s\"}\n", e.what());
        return 4;
    }
    return 0;
}
