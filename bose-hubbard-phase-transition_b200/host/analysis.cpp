// analysis.cpp -- the (J, U, mu) sweep driver on top of the C ABI.
//
// Mirrors Analysis::exact_parameters / calculate_and_save (reference src/analysis.cpp:207-427): same range
// plumbing (including the parameter mix-up of :242-256), same grid-size arithmetic (:281-282),
// same result indexing (:341), same variance-restart rule (:364-380) and the same phase.txt format (:269-274,
// :384-387).  Grid points are independent, so instead of an OpenMP team the points are pulled from a shared
// counter by one host thread per GPU, each owning a bh_ctx (no collective on the data path).
#include "analysis.hpp"

#include <atomic>
#include <cmath>
#include <cuda_runtime_api.h>
#include <exception>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <mutex>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <thread>

#include "../../include/bh_b200.h"
#include "neighbours.hpp"
#include "resource.hpp"

namespace {

struct Grid {
    std::string fixed;
    double fixed_value;
    double p1_min, p1_max, p2_min, p2_max, step1, step2;
    int num1, num2;
};

// coefficient triple (cJ, cU, cmu) of a grid point, by sweep mode (src/analysis.cpp:245-256)
void coefficients(const Grid& g, double p1, double p2, double& cJ, double& cU, double& cmu)
{
    if (g.fixed == "J") {          // H = JH*J + UH*p1 + uH*p2
        cJ = g.fixed_value; cU = p1; cmu = p2;
    } else if (g.fixed == "U") {   // H = UH*U + JH*p1 + uH*p2
        cU = g.fixed_value; cJ = p1; cmu = p2;
    } else {                       // H = uH*mu + JH*p1 + UH*p2
        cmu = g.fixed_value; cJ = p1; cU = p2;
    }
}

std::vector<Analysis::SweepPoint> calculate_and_save(int m, int n, const std::vector<std::vector<int>>& nei, const Grid& g,
                                                     const Analysis::ExactOptions& opt)
{
    std::ofstream file(opt.output);
    file << g.fixed << " " << g.fixed_value << std::endl;

    int nb_eigen = opt.nb_eigen;
    const int total = g.num1 * g.num2;
    std::vector<Analysis::SweepPoint> res(total, Analysis::SweepPoint{0, 0, 0, 0, 0});

    // neighbour list in the C ABI's CSR form
    std::vector<int> nbr_ptr(m + 1, 0), nbr_idx;
    for (int i = 0; i < m; ++i) {
        nbr_ptr[i + 1] = nbr_ptr[i] + static_cast<int>(nei[i].size());
        nbr_idx.insert(nbr_idx.end(), nei[i].begin(), nei[i].end());
    }
    if (nbr_idx.empty()) nbr_idx.push_back(0);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        throw std::runtime_error("no CUDA device: the exact path has no CPU fallback");
    const int ngpu = (opt.gpus > 0) ? std::min(opt.gpus, ndev) : ndev;

    // Optional: several grid points concurrently per GPU, each on its own context (own stream and workspace).
    // Measured on B200 (m=n=8 and 10, 121 points): no gain -- cooperative launches of different streams serialise
    // and the host threads contend for the driver -- so the default is one context per GPU.
    int per_gpu = std::max(1, opt.contexts_per_gpu);
    const int nworkers = ngpu * per_gpu;
    std::vector<bh_ctx*> ctxs(nworkers, nullptr);
    for (int d = 0; d < nworkers; ++d) {
        if (bh_ctx_create(d % ngpu, &ctxs[d]) != BH_OK) throw std::runtime_error(bh_last_error(nullptr));
        bh_ctx_set_batch(ctxs[d], std::max(1, std::min(4, opt.batch)));
        if (bh_setup(ctxs[d], m, n, nbr_ptr.data(), nbr_idx.data()) != BH_OK) {
            std::string msg = bh_last_error(ctxs[d]);
            for (bh_ctx* c : ctxs) bh_ctx_destroy(c);
            throw std::runtime_error(msg);
        }
    }

    const double variance_threshold_percent = 1e-8;  // src/analysis.cpp:298
    const int bar_width = 100;
    std::exception_ptr failure;
    std::mutex mtx;

    // In the -f J and -f U modes the second swept parameter multiplies uH = -N*I (src/analysis.cpp:245-251): it only
    // shifts the spectrum, eigenvectors and all three output columns are unchanged.  With
    // opt.reuse_shift the row j = 0 is solved and copied to the other j (exact identity, opt-in, off by default).
    // Checkpoint / resume (the reference writes phase.txt only after the whole grid, src/analysis.cpp:384-387): with
    // opt.resume every finished point is appended to "<output>.partial" (index + the five columns, 17 significant
    // digits) and points found there are not recomputed.  The file is tied to the sweep by a header line.
    const std::string partial_path = opt.output + ".partial";
    std::vector<char> have(total, 0);
    std::ofstream partial;
    // the header ties the checkpoint to the sweep AND to the current nb_eigen: rows written before a variance restart
    // (nb_eigen += 5, the whole grid repeated) must never be mixed with rows of the new round
    auto partial_header = [&](int nbe) {
        std::ostringstream hdr;
        hdr << "# m " << m << " n " << n << " fixed " << g.fixed << " " << std::setprecision(17) << g.fixed_value << " grid " << g.p1_min << " "
            << g.p2_min << " " << g.step1 << " " << g.num1 << " " << g.num2 << " nb_eigen " << nbe;
        return hdr.str();
    };
    if (opt.resume) {
        // a run killed after a variance restart left a header with a larger nb_eigen: resume in that round
        {
            std::ifstream probe(partial_path);
            std::string first;
            if (std::getline(probe, first)) {
                for (int nbe = nb_eigen + 5; nbe <= nb_eigen + 100; nbe += 5)
                    if (first == partial_header(nbe)) {
                        nb_eigen = nbe;
                        break;
                    }
            }
        }
        std::ostringstream hdr;
        hdr << partial_header(nb_eigen);
        std::ifstream in(partial_path);
        std::string line;
        bool ok = static_cast<bool>(std::getline(in, line)) && line == hdr.str();
        while (ok && std::getline(in, line)) {
            std::istringstream ls(line);
            int index;
            Analysis::SweepPoint p;
            if (ls >> index >> p.param1 >> p.param2 >> p.gap_ratio >> p.condensate_fraction >> p.coherence && index >= 0 && index < total) {
                res[index] = p;
                have[index] = 1;
            }
        }
        in.close();
        partial.open(partial_path, ok ? std::ios::app : std::ios::trunc);
        if (!ok) partial << hdr.str() << std::endl;
        partial << std::setprecision(17);
    }
    const int kernel = opt.kernel >= 0 ? opt.kernel : (opt.lx > 0 ? BH_HV_STORED : BH_HV_MATRIX_FREE);
    const bool shift_rows = opt.reuse_shift && g.fixed != "u";
    const int ntasks = shift_rows ? g.num1 : total;
    while (true) {
        std::atomic<int> next(0), done(0);
        auto worker = [&](int d) {
            try {
                const int B = std::max(1, std::min(4, opt.batch));
                int64_t dim = 0;
                bh_dimension(m, n, &dim);
                // a chunk of tasks per call: bh_points keeps B solves in lockstep (shared H.v launches, same results) and
                // refills a finished solve from the chunk; small systems (D <= 100 000) are solved one CTA per grid point,
                // so they are handed over in chunks large enough to fill the SMs of the GPU
                const int chunk = B > 1 ? ((dim <= 100000 && kernel == BH_HV_MATRIX_FREE) ? 160 : std::min(16, 4 * B)) : 1;
                std::vector<int> ts(chunk);
                std::vector<double> cJ(chunk), cU(chunk), cmu(chunk), p1s(chunk), out3(3 * (size_t)chunk);
                for (;;) {
                    int nt = 0;
                    while (nt < chunk) {
                        const int t = next.fetch_add(1);
                        if (t >= ntasks) break;
                        const int i = shift_rows ? t : t / g.num2, j = shift_rows ? 0 : t % g.num2;
                        // resume: every output row this task would produce is already there
                        bool all = opt.resume;
                        for (int jj = j; all && jj < (shift_rows ? g.num2 : j + 1); ++jj) {
                            const int index = i * g.num1 + jj;
                            all = index >= 0 && index < total && have[index];
                        }
                        if (all) {
                            std::lock_guard<std::mutex> lk(mtx);
                            done += shift_rows ? g.num2 : 1;
                            continue;
                        }
                        p1s[nt] = g.p1_min + i * g.step1;
                        coefficients(g, p1s[nt], g.p2_min + j * g.step2, cJ[nt], cU[nt], cmu[nt]);
                        ts[nt++] = t;
                    }
                    if (nt == 0) break;
                    const int rc = nt == 1 ? bh_point(ctxs[d], cJ[0], cU[0], cmu[0], nb_eigen, kernel, out3.data(), nullptr, nullptr, nullptr)
                                           : bh_points(ctxs[d], cJ.data(), cU.data(), cmu.data(), nt, nb_eigen, kernel, out3.data(), nullptr);
                    if (rc == BH_ERR_ARG) throw std::invalid_argument(bh_last_error(ctxs[d]));
                    if (rc != BH_OK) throw std::runtime_error(bh_last_error(ctxs[d]));
                    std::lock_guard<std::mutex> lk(mtx);
                    for (int q = 0; q < nt; ++q) {
                        const int t = ts[q];
                        const int i = shift_rows ? t : t / g.num2, j = shift_rows ? 0 : t % g.num2;
                        const double* o3 = out3.data() + 3 * q;
                        for (int jj = j; jj < (shift_rows ? g.num2 : j + 1); ++jj) {
                            const int index = i * g.num1 + jj;  // src/analysis.cpp:341 (sic)
                            if (index >= 0 && index < total) {
                                res[index] = Analysis::SweepPoint{p1s[q], g.p2_min + jj * g.step2, o3[0], o3[1], o3[2]};
                                if (opt.resume)
                                    partial << index << " " << res[index].param1 << " " << res[index].param2 << " " << o3[0] << " " << o3[1]
                                            << " " << o3[2] << std::endl;
                            }
                            ++done;
                        }
                    }
                    const int c = done.load();
                    if (opt.progress) {
                        const int progress = (c * bar_width) / total;
                        std::cout << "\rProgress: [" << std::string(progress, '#') << std::string(bar_width - progress, ' ') << "] "
                                  << std::setw(3) << (c * 100) / total << "% " << std::flush;
                    }
                }
            } catch (...) {
                std::lock_guard<std::mutex> lk(mtx);
                if (!failure) failure = std::current_exception();
                next.store(total);
            }
        };
        std::vector<std::thread> team;
        for (int d = 0; d < nworkers; ++d) team.emplace_back(worker, d);
        for (auto& th : team) th.join();
        if (failure) break;

        // src/analysis.cpp:364-380
        double mean = 0.0;
        for (const auto& p : res) mean += p.gap_ratio;
        mean /= res.size();
        double variance = 0.0;
        for (const auto& p : res) variance += (p.gap_ratio - mean) * (p.gap_ratio - mean);
        variance /= res.size();
        if (variance > variance_threshold_percent * mean) break;
        nb_eigen += 5;
        std::fill(have.begin(), have.end(), 0);  // the whole grid is repeated with more eigenvalues
        if (opt.resume) {  // new round, new checkpoint: the rows of the previous nb_eigen are dropped
            partial.close();
            partial.open(partial_path, std::ios::trunc);
            partial << partial_header(nb_eigen) << std::endl << std::setprecision(17);
        }
    }
    for (bh_ctx* c : ctxs) bh_ctx_destroy(c);
    if (failure) std::rethrow_exception(failure);

    for (const auto& p : res)
        file << p.param1 << " " << p.param2 << " " << p.gap_ratio << " " << p.condensate_fraction << " " << p.coherence << std::endl;
    file.close();
    return res;
}

}  // namespace

std::vector<Analysis::SweepPoint> Analysis::exact_parameters(int m, int n, double J, double U, double mu, double s, double r,
                                                             std::string fixed_param, const ExactOptions& opt)
{
    const double eps = std::numeric_limits<double>::epsilon();
    if (std::abs(J - 0.0) < eps && std::abs(U - 0.0) < eps && std::abs(mu - 0.0) < eps) {
        std::cerr << "Error: At least one of the parameters J, U, mu must be different from zero." << std::endl;
        return {};
    }
    Resource::timer();

    Neighbours neighbours(m);
    if (opt.lx > 0)
        neighbours.rect_neighbours(opt.lx, opt.ly > 0 ? opt.ly : 1, opt.lz > 0 ? opt.lz : 1, opt.closed);
    else
        neighbours.chain_neighbours(opt.closed);
    const std::vector<std::vector<int>> nei = neighbours.getNeighbours();

    // src/analysis.cpp:242-256
    const double J_min = J, J_max = J + r, mu_min = mu, mu_max = mu + r, U_min = U, U_max = U + r;
    Grid g;
    g.fixed = fixed_param;
    g.step1 = g.step2 = s;
    if (fixed_param == "J") {
        g.fixed_value = J; g.p1_min = J_min; g.p1_max = J_max; g.p2_min = mu_min; g.p2_max = mu_max;
    } else if (fixed_param == "U") {
        g.fixed_value = U; g.p1_min = J_min; g.p1_max = J_max; g.p2_min = U_min; g.p2_max = U_max;
    } else {
        g.fixed_value = mu; g.p1_min = J_min; g.p1_max = J_max; g.p2_min = mu_min; g.p2_max = mu_max;
    }
    // src/analysis.cpp:281-282, evaluated in double exactly as written
    g.num1 = static_cast<int>((g.p1_max - g.p1_min) / g.step1) + 1;
    g.num2 = static_cast<int>((g.p2_max - g.p2_min) / g.step2) + 1;

    std::vector<SweepPoint> rows = calculate_and_save(m, n, nei, g, opt);

    std::cout << std::endl;
    Resource::timer();
    Resource::get_memory_usage(true);
    return rows;
}

void Analysis::exact_parameters(int m, int n, double J, double U, double mu, double s, double r, std::string fixed_param)
{
    ExactOptions opt;
    exact_parameters(m, n, J, U, mu, s, r, fixed_param, opt);
}

void Analysis::mean_field_parameters(int, int)
{
    std::cerr << "The mean-field calculation is outside the accelerated exact-diagonalisation path; "
                 "use the reference program for -t mean." << std::endl;
}
