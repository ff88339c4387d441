// analysis.hpp -- drop-in for the exact-diagonalisation entry point of the reference's namespace Analysis
// (include/analysis.hpp:99, src/analysis.cpp:207-427).  Eigen-free: the sweep runs entirely behind the C ABI.
#pragma once

#include <string>
#include <vector>

namespace Analysis {

// Options the reference hard-codes (src/analysis.cpp:219-220, :269, :277).  The defaults reproduce it.
struct ExactOptions {
    int gpus = 0;                  // GPUs to shard the grid over (0 = every visible device)
    bool resume = false;           // append finished points to <output>.partial and skip them when restarted
    bool reuse_shift = false;      // -f J / -f U: the second parameter only shifts the spectrum; solve each row once
    int contexts_per_gpu = 1;      // concurrent grid points per GPU (each on its own context)
    int kernel = -1;               // BH_HV_STORED (0), BH_HV_MATRIX_FREE (1); -1 = matrix-free for chains (faster, lockstep-capable), stored for other lattices
    int batch = 4;                 // grid points solved in lockstep per context (bh_ctx_set_batch; effective for chains + matrix-free)
    int lx = 0, ly = 0, lz = 0;    // lx*ly*lz == m selects a periodic box; all zero = closed chain (reference)
    bool closed = true;
    std::string output = "phase.txt";
    bool progress = true;          // the reference's "\rProgress: [####   ] NN%" bar
    int nb_eigen = 20;             // src/analysis.cpp:277
};

struct SweepPoint {
    double param1, param2, gap_ratio, condensate_fraction, coherence;
};

// Reference signature (include/analysis.hpp:99): writes phase.txt in the current directory.
void exact_parameters(int m, int n, double J, double U, double mu, double s, double r, std::string fixed_param);
// Same computation with the options above; returns the rows written (in file order).
std::vector<SweepPoint> exact_parameters(int m, int n, double J, double U, double mu, double s, double r,
                                         std::string fixed_param, const ExactOptions& opt);
// The mean-field path (src/analysis.cpp:58-178) is outside the accelerated scope: reports that and returns.
void mean_field_parameters(int n, int precision);

}  // namespace Analysis
