// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin command-line driver around the reference's own translation units, compiled
// by oracle/Makefile from /root/reference (patched P1..P6 by oracle/patch_ref.py,
// see SURVEY.md section 8c) into oracle/_ref/ref_harness.  It is used
//   * to pin oracle/bh_oracle.cpp (the CPU restatement) against the real reference,
//   * to generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   * as the CPU baseline (`bench.py --impl reference`, cpu_baseline.kind = "reference").
//
// The reference's `Analysis::*` helpers are namespace-scope statics, so the
// patched .cpp files are #included here as one unity translation unit.
//
// With -DBH_REF_UNPATCHED only `basis` is available and it runs the *unpatched*
// BH::fixed_set_basis (well defined; its scatter order is a parity target).
//
// Every array is dumped raw (little endian) to <out>.<name>.{f64,i32}.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include <omp.h>

#ifdef BH_REF_UNPATCHED
#include "hamiltonian_unpatched.cpp"
#else
#include "hamiltonian.cpp"
#include "operator.cpp"
#include "neighbours.cpp"
#include "resource.cpp"
#include "analysis.cpp"
#endif

static void dump(const std::string& out, const char* name, const double* p, size_t n)
{
    std::string f = out + "." + name + ".f64";
    FILE* fp = fopen(f.c_str(), "wb");
    if (!fp) { perror(f.c_str()); exit(2); }
    fwrite(p, sizeof(double), n, fp);
    fclose(fp);
}
static void dump(const std::string& out, const char* name, const int* p, size_t n)
{
    std::string f = out + "." + name + ".i32";
    FILE* fp = fopen(f.c_str(), "wb");
    if (!fp) { perror(f.c_str()); exit(2); }
    fwrite(p, sizeof(int), n, fp);
    fclose(fp);
}
static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#ifndef BH_REF_UNPATCHED
// lattice spec: "chain" (reference Neighbours::chain_neighbours, closed) or "rect:LX:LY"
// (periodic LX x LY rectangle, site = y*LX + x, order left,right,up,down -- the
// reference cannot produce it, SURVEY.md D7, so the list is built here and handed
// to BH::fixed_bosons_hamiltonian which accepts any neighbour list).
static std::vector<std::vector<int>> lattice(const std::string& spec, int m)
{
    if (spec == "chain") {
        Neighbours nb(m);
        nb.chain_neighbours();
        return nb.getNeighbours();
    }
    if (spec == "openchain") {
        Neighbours nb(m);
        nb.chain_neighbours(false);
        return nb.getNeighbours();
    }
    int lx = 0, ly = 0;
    if (sscanf(spec.c_str(), "rect:%d:%d", &lx, &ly) == 2 && lx * ly == m) {
        std::vector<std::vector<int>> nei(m);
        for (int y = 0; y < ly; ++y)
            for (int x = 0; x < lx; ++x) {
                int s = y * lx + x;
                nei[s].push_back(y * lx + (x + lx - 1) % lx);
                nei[s].push_back(y * lx + (x + 1) % lx);
                nei[s].push_back(((y + ly - 1) % ly) * lx + x);
                nei[s].push_back(((y + 1) % ly) * lx + x);
            }
        return nei;
    }
    fprintf(stderr, "bad lattice spec %s\n", spec.c_str());
    exit(2);
}

static void dump_csc(const std::string& out, Eigen::SparseMatrix<double>& H)
{
    H.makeCompressed();
    dump(out, "outer", H.outerIndexPtr(), H.outerSize() + 1);
    dump(out, "inner", H.innerIndexPtr(), H.nonZeros());
    dump(out, "val", H.valuePtr(), H.nonZeros());
}

struct Terms {
    Eigen::VectorXd tags;
    Eigen::MatrixXd basis;
    Eigen::SparseMatrix<double> JH, UH, uH;
};
static Terms build_terms(int m, int n, const std::string& lat)
{
    Terms t;
    auto nei = lattice(lat, m);
    auto tb = BH::fixed_set_basis(m, n);
    t.tags = tb.first;
    t.basis = tb.second;
    t.JH = BH::fixed_bosons_hamiltonian(nei, t.basis, t.tags, m, n, 1, 0, 0);
    t.UH = BH::fixed_bosons_hamiltonian(nei, t.basis, t.tags, m, n, 0, 1, 0);
    t.uH = BH::fixed_bosons_hamiltonian(nei, t.basis, t.tags, m, n, 0, 0, 1);
    return t;
}

// One grid point exactly as the body of the reference sweep loop does it
// (src/analysis.cpp:311-337): H, IRLM_eigen, gap ratios, SPDM, condensate fraction, coherence.
static void one_point(const Terms& t, const Eigen::SparseMatrix<double>& Hfixed,
                      const Eigen::SparseMatrix<double>& H1, const Eigen::SparseMatrix<double>& H2,
                      double p1, double p2, int nb_eigen, double* out5, double* evals, double* rho)
{
    Eigen::SparseMatrix<double> H = Hfixed + H1 * p1 + H2 * p2;
    Eigen::MatrixXcd eigenvectors;
    Eigen::VectorXcd eigenvalues = Op::IRLM_eigen(H, nb_eigen, eigenvectors);
    Eigen::VectorXd vec_ratios = Analysis::gap_ratios(eigenvalues, nb_eigen);
    double gap_ratio = vec_ratios.size() > 0 ? vec_ratios.sum() / vec_ratios.size() : 0.0;
    Eigen::MatrixXcd spdm = Analysis::SPDM(t.basis, t.tags, eigenvectors);
    Eigen::EigenSolver<Eigen::MatrixXd> solver(spdm.real());
    double cf = std::abs(std::max_element(solver.eigenvalues().begin(), solver.eigenvalues().end(),
                                          [](const std::complex<double>& a, const std::complex<double>& b) {
                                              return std::abs(a) < std::abs(b);
                                          })->real() / spdm.trace());
    double K = Analysis::coherence(spdm);
    out5[0] = p1; out5[1] = p2; out5[2] = gap_ratio; out5[3] = cf; out5[4] = K;
    if (evals) for (int i = 0; i < nb_eigen; ++i) evals[i] = eigenvalues[i].real();
    if (rho) {
        int m = (int)spdm.rows();
        for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) rho[i + j * m] = spdm(i, j).real();
    }
}
#endif

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: ref_harness <cmd> ...\n"); return 2; }
    std::string cmd = argv[1];
    if (cmd == "basis") {
        // basis m n out
        int m = atoi(argv[2]), n = atoi(argv[3]);
        std::string out = argv[4];
        double t0 = now();
        auto tb = BH::fixed_set_basis(m, n);
        double t1 = now();
        dump(out, "tags", tb.first.data(), tb.first.size());
        dump(out, "basis", tb.second.data(), tb.second.size());
        printf("{\"D\": %ld, \"seconds\": %.6f}\n", (long)tb.first.size(), t1 - t0);
        return 0;
    }
#ifndef BH_REF_UNPATCHED
    if (cmd == "csc") {
        // csc m n term lattice out      term in J,U,u : the three calls of src/analysis.cpp:234-236
        int m = atoi(argv[2]), n = atoi(argv[3]);
        std::string term = argv[4], lat = argv[5], out = argv[6];
        auto nei = lattice(lat, m);
        auto tb = BH::fixed_set_basis(m, n);
        double t0 = now();
        Eigen::SparseMatrix<double> H = BH::fixed_bosons_hamiltonian(
            nei, tb.second, tb.first, m, n, term == "J" ? 1 : 0, term == "U" ? 1 : 0, term == "u" ? 1 : 0);
        double t1 = now();
        dump_csc(out, H);
        printf("{\"D\": %ld, \"nnz\": %ld, \"seconds\": %.6f}\n", (long)H.rows(), (long)H.nonZeros(), t1 - t0);
        return 0;
    }
    if (cmd == "hsum") {
        // hsum m n cJ cU cu lattice out : H = JH*cJ + UH*cU + uH*cu built the way analysis.cpp:311 does (-f J mode)
        int m = atoi(argv[2]), n = atoi(argv[3]);
        double cJ = atof(argv[4]), cU = atof(argv[5]), cu = atof(argv[6]);
        std::string lat = argv[7], out = argv[8];
        Terms t = build_terms(m, n, lat);
        Eigen::SparseMatrix<double> Hf = t.JH * cJ;
        Eigen::SparseMatrix<double> H = Hf + t.UH * cU + t.uH * cu;
        dump_csc(out, H);
        printf("{\"D\": %ld, \"nnz\": %ld}\n", (long)H.rows(), (long)H.nonZeros());
        return 0;
    }
    if (cmd == "eigs") {
        // eigs m n cJ cU cu nev lattice out : one grid point; dumps evals, rho (m*m col-major), out5
        int m = atoi(argv[2]), n = atoi(argv[3]);
        double cJ = atof(argv[4]), cU = atof(argv[5]), cu = atof(argv[6]);
        int nev = atoi(argv[7]);
        std::string lat = argv[8], out = argv[9];
        Terms t = build_terms(m, n, lat);
        Eigen::SparseMatrix<double> Hf = t.JH * cJ;
        std::vector<double> ev(nev), rho(m * m), o5(5);
        double t0 = now();
        one_point(t, Hf, t.UH, t.uH, cU, cu, nev, o5.data(), ev.data(), rho.data());
        double t1 = now();
        dump(out, "evals", ev.data(), ev.size());
        dump(out, "rho", rho.data(), rho.size());
        dump(out, "out5", o5.data(), 5);
        // matvec / restart counts of the same solve (direct Spectra call with the same settings)
        Eigen::SparseMatrix<double> H = Hf + t.UH * cU + t.uH * cu;
        Spectra::SparseGenMatProd<double> op(H);
        Spectra::GenEigsSolver<Spectra::SparseGenMatProd<double>> eigs(op, nev, 2 * nev + 1);
        eigs.init();
        eigs.compute(Spectra::SortRule::SmallestReal, 1000, 1e-10, Spectra::SortRule::SmallestReal);
        printf("{\"D\": %ld, \"seconds\": %.6f, \"nmatvec\": %ld, \"nrestart\": %ld}\n", (long)H.rows(), t1 - t0,
               (long)eigs.num_operations(), (long)eigs.num_iterations());
        return 0;
    }
    if (cmd == "hv") {
        // hv m n cJ cU cu reps lattice out : y = H x through Spectra's MatOp (SparseGenMatProd.h:81-86),
        // x = Spectra LCG(seed 0) uniform(-0.5,0.5); prints seconds per H.v
        int m = atoi(argv[2]), n = atoi(argv[3]);
        double cJ = atof(argv[4]), cU = atof(argv[5]), cu = atof(argv[6]);
        int reps = atoi(argv[7]);
        std::string lat = argv[8], out = argv[9];
        Terms t = build_terms(m, n, lat);
        Eigen::SparseMatrix<double> Hf = t.JH * cJ;
        Eigen::SparseMatrix<double> H = Hf + t.UH * cU + t.uH * cu;
        Spectra::SparseGenMatProd<double> op(H);
        Spectra::SimpleRandom<double> rng(0);
        Eigen::VectorXd x = rng.random_vec(H.rows());
        Eigen::VectorXd y(H.rows());
        op.perform_op(x.data(), y.data());
        double t0 = now();
        for (int r = 0; r < reps; ++r) op.perform_op(x.data(), y.data());
        double t1 = now();
        if (out != "-") {
            dump(out, "x", x.data(), x.size());
            dump(out, "y", y.data(), y.size());
        }
        printf("{\"D\": %ld, \"nnz\": %ld, \"seconds_per_hv\": %.9f, \"reps\": %d}\n", (long)H.rows(), (long)H.nonZeros(),
               (t1 - t0) / (reps > 0 ? reps : 1), reps);
        return 0;
    }
    if (cmd == "thermal") {
        // thermal m n cJ cU cu T out : the finite-temperature branch of the sweep body (src/analysis.cpp:321-323, 474-494;
        // dead in the reference, whose temperature is the constant 0): density_matrix(eigenvalues, eigenvectors, T) of the
        // 20 Ritz pairs -- a dense D x D matrix -- and the two scalars the loop derives from it (:331-337).
        int m = atoi(argv[2]), n = atoi(argv[3]);
        double cJ = atof(argv[4]), cU = atof(argv[5]), cu = atof(argv[6]), T = atof(argv[7]);
        std::string out = argv[8];
        Terms t = build_terms(m, n, "chain");
        Eigen::SparseMatrix<double> Hf = t.JH * cJ;
        Eigen::SparseMatrix<double> H = Hf + t.UH * cU + t.uH * cu;
        Eigen::MatrixXcd eigenvectors;
        Eigen::VectorXcd eigenvalues = Op::IRLM_eigen(H, 20, eigenvectors);
        Eigen::MatrixXcd dm = Analysis::density_matrix(eigenvalues, eigenvectors, T);
        Eigen::MatrixXd re = dm.real();
        Eigen::EigenSolver<Eigen::MatrixXd> solver(re);
        double cf = std::abs(std::max_element(solver.eigenvalues().begin(), solver.eigenvalues().end(),
                                              [](const std::complex<double>& a, const std::complex<double>& b) {
                                                  return std::abs(a) < std::abs(b);
                                              })->real() / dm.trace());
        double K = Analysis::coherence(dm);
        std::vector<double> ev(20), o2{cf, K};
        for (int i = 0; i < 20; ++i) ev[i] = eigenvalues[i].real();
        dump(out, "dm", re.data(), re.size());
        dump(out, "evals", ev.data(), ev.size());
        dump(out, "out2", o2.data(), o2.size());
        printf("{\"D\": %ld}\n", (long)H.rows());
        return 0;
    }
    if (cmd == "maxbasis") {
        // maxbasis m n out : BH::max_set_basis (src/hamiltonian.cpp:152-166), all boson numbers 1..n, tag-sorted
        int m = atoi(argv[2]), n = atoi(argv[3]);
        std::string out = argv[4];
        auto tb = BH::max_set_basis(m, n);
        dump(out, "tags", tb.first.data(), tb.first.size());
        dump(out, "basis", tb.second.data(), tb.second.size());
        printf("{\"D\": %ld}\n", (long)tb.first.size());
        return 0;
    }
    if (cmd == "maxham") {
        // maxham m nmin nmax J U u lattice out : BH::max_bosons_hamiltonian (src/hamiltonian.cpp:260-288)
        int m = atoi(argv[2]), nmin = atoi(argv[3]), nmax = atoi(argv[4]);
        double J = atof(argv[5]), U = atof(argv[6]), mu = atof(argv[7]);
        std::string lat = argv[8], out = argv[9];
        auto nei = lattice(lat, m);
        Eigen::SparseMatrix<double> H = BH::max_bosons_hamiltonian(nei, m, nmin, nmax, J, U, mu);
        dump_csc(out, H);
        printf("{\"D\": %ld, \"nnz\": %ld}\n", (long)H.rows(), (long)H.nonZeros());
        return 0;
    }
    if (cmd == "partial") {
        // partial m n cJ cU cu maxit threads lattice
        // Bounded sample of one grid point for the CPU baseline at sizes where a full solve takes minutes:
        // `threads` concurrent copies (one per core, as the sweep's OpenMP loop would run them) of the
        // reference solver call of src/operator.cpp:22-26 on the same H, stopped after `maxit` restarts.
        int m = atoi(argv[2]), n = atoi(argv[3]);
        double cJ = atof(argv[4]), cU = atof(argv[5]), cu = atof(argv[6]);
        int maxit = atoi(argv[7]), threads = atoi(argv[8]);
        std::string lat = argv[9];
        double tb0 = now();
        Terms t = build_terms(m, n, lat);
        double tb1 = now();
        Eigen::SparseMatrix<double> Hf = t.JH * cJ;
        Eigen::SparseMatrix<double> H = Hf + t.UH * cU + t.uH * cu;
        if (threads < 1) threads = omp_get_max_threads();
        std::vector<long> ops(threads), its(threads), conv(threads);
        double t0 = now();
#pragma omp parallel for num_threads(threads) schedule(static, 1)
        for (int c = 0; c < threads; ++c) {
            Spectra::SparseGenMatProd<double> op(H);
            Spectra::GenEigsSolver<Spectra::SparseGenMatProd<double>> eigs(op, 20, 41);
            eigs.init();
            int nconv = eigs.compute(Spectra::SortRule::SmallestReal, maxit, 1e-10, Spectra::SortRule::SmallestReal);
            ops[c] = eigs.num_operations();
            its[c] = eigs.num_iterations();
            conv[c] = nconv;
        }
        double t1 = now();
        printf("{\"D\": %ld, \"nnz\": %ld, \"threads\": %d, \"seconds\": %.6f, \"setup_seconds\": %.6f, \"nmatvec\": %ld, "
               "\"nrestart\": %ld, \"nconv\": %ld}\n",
               (long)H.rows(), (long)H.nonZeros(), threads, t1 - t0, tb1 - tb0, ops[0], its[0], conv[0]);
        return 0;
    }
    if (cmd == "points") {
        // points m n fixed cfix p1min p2min step n1 n2 threads lattice out
        // A (sub)grid of the sweep, body identical to analysis.cpp:302-343, OpenMP over points with
        // the thread heuristic of resource.cpp:77-83 BYPASSED (it collapses to 1 thread).
        int m = atoi(argv[2]), n = atoi(argv[3]);
        std::string fixed = argv[4];
        double cfix = atof(argv[5]), p1min = atof(argv[6]), p2min = atof(argv[7]), step = atof(argv[8]);
        int n1 = atoi(argv[9]), n2 = atoi(argv[10]), threads = atoi(argv[11]);
        std::string lat = argv[12], out = argv[13];
        double tb0 = now();
        Terms t = build_terms(m, n, lat);
        double tb1 = now();
        Eigen::SparseMatrix<double> Hf, H1, H2;
        if (fixed == "J") { Hf = t.JH * cfix; H1 = t.UH; H2 = t.uH; }
        else if (fixed == "U") { Hf = t.UH * cfix; H1 = t.JH; H2 = t.uH; }
        else { Hf = t.uH * cfix; H1 = t.JH; H2 = t.UH; }
        std::vector<double> res(5 * (size_t)n1 * n2), evs(20 * (size_t)n1 * n2);
        if (threads > 0) omp_set_num_threads(threads);
        double t0 = now();
#pragma omp parallel for collapse(2) schedule(dynamic)
        for (int i = 0; i < n1; ++i)
            for (int j = 0; j < n2; ++j) {
                size_t idx = (size_t)i * n2 + j;
                one_point(t, Hf, H1, H2, p1min + i * step, p2min + j * step, 20, &res[5 * idx], &evs[20 * idx], nullptr);
            }
        double t1 = now();
        if (out != "-") {
            dump(out, "out5", res.data(), res.size());
            dump(out, "evals", evs.data(), evs.size());
        }
        printf("{\"D\": %ld, \"points\": %d, \"seconds\": %.6f, \"setup_seconds\": %.6f, \"threads\": %d}\n", (long)t.tags.size(),
               n1 * n2, t1 - t0, tb1 - tb0, threads > 0 ? threads : omp_get_max_threads());
        return 0;
    }
#endif
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}
