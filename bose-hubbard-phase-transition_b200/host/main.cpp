// main.cpp -- command line of the reference program (src/main.cpp:90-204) on the B200 path.
//
// Same options: -m -n -J -U -u -r -s -f -t [-i -e -h] and their long names; same validation messages and exit
// codes.  Additions (all optional, defaults = reference behaviour): -g/--gpus N, -l/--lattice chain|LXxLY[xLZ],
// -k/--kernel stored|free, -o/--output FILE, --no-plot.  After an exact run the reference executes
// `python3 plot.py`; this does the same when plot.py exists in the working directory.
#include <getopt.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "analysis.hpp"

static void print_usage()
{
    std::cout << "Usage: program [options]\n"
              << "Options:\n"
              << "  -m, --sites       Number of sites\n"
              << "  -n, --bosons      Number of bosons\n"
              << "  -J, --hopping     Hopping parameter\n"
              << "  -U, --interaction On-site interaction\n"
              << "  -u, --potential   Chemical potential\n"
              << "  -r, --range     Range for varying parameters (if range is the same for each)\n"
              << "  -s, --step      Step for varying parameters (with s < r)\n"
              << "  -f --fixed      Fixed parameter (J, U or u) \n"
              << "  -t, --type      Type of calculation (exact or mean; 'mean' is NOT part of the accelerated path: exit code 2)\n"
              << "  -i, --iterations  Number of iterations over the parameters in the mean-field approximation\n"
              << "  -e, --epsilon  Threshold for convergence in the mean-field approximation\n"
              << "  -g, --gpus      GPUs to shard the sweep over (default: all visible)\n"
              << "  -l, --lattice   chain (default) or LXxLY[xLZ] periodic box\n"
              << "  -k, --kernel    stored or free (matrix-free H.v); default: free for chains, stored for --lattice\n"
              << "  -o, --output    Output file (default phase.txt)\n"
              << "  -c, --concurrent Grid points solved concurrently per GPU (default 1)\n"
              << "      --no-plot   Do not run plot.py afterwards\n"
              << "      --resume    Checkpoint finished points in <output>.partial and skip them when restarted\n"
              << "      --batch N   Grid points solved in lockstep per GPU, sharing their H.v launches (1..4, default 4)\n"
              << "      --reuse-shift  -f J / -f U: the chemical potential only shifts the spectrum; solve each row once\n";
}

int main(int argc, char* argv[])
{
    int m = 0, n = 0, it = 0, eps = 0;
    double J = 0, U = 0, mu = 0, s = 0, r = 0;
    std::string fixed_param, calc_type;
    Analysis::ExactOptions opt;
    bool plot = true;

    const char* const short_opts = "m:n:J:U:u:r:s:f:t:i:e:g:l:k:o:c:h";
    const option long_opts[] = {{"sites", required_argument, nullptr, 'm'},      {"bosons", required_argument, nullptr, 'n'},
                                {"hopping", required_argument, nullptr, 'J'},    {"interaction", required_argument, nullptr, 'U'},
                                {"potential", required_argument, nullptr, 'u'},  {"range", required_argument, nullptr, 'r'},
                                {"step", required_argument, nullptr, 's'},       {"fixed", required_argument, nullptr, 'f'},
                                {"type", required_argument, nullptr, 't'},       {"iterations", required_argument, nullptr, 'i'},
                                {"epsilon", required_argument, nullptr, 'e'},    {"gpus", required_argument, nullptr, 'g'},
                                {"lattice", required_argument, nullptr, 'l'},    {"kernel", required_argument, nullptr, 'k'},
                                {"output", required_argument, nullptr, 'o'},     {"concurrent", required_argument, nullptr, 'c'},     {"no-plot", no_argument, nullptr, 1000},         {"reuse-shift", no_argument, nullptr, 1001},      {"resume", no_argument, nullptr, 1002},         {"batch", required_argument, nullptr, 1003},
                                {"help", no_argument, nullptr, 'h'},             {nullptr, no_argument, nullptr, 0}};
    while (true) {
        const int o = getopt_long(argc, argv, short_opts, long_opts, nullptr);
        if (o == -1) break;
        switch (o) {
            case 'm': m = std::stoi(optarg); break;
            case 'n': n = std::stoi(optarg); break;
            case 'J': J = std::stod(optarg); break;
            case 'U': U = std::stod(optarg); break;
            case 'u': mu = std::stod(optarg); break;
            case 'r': r = std::stod(optarg); break;
            case 's': s = std::stod(optarg); break;
            case 'f': fixed_param = optarg; break;
            case 't': calc_type = optarg; break;
            case 'i': it = static_cast<int>(std::stod(optarg)); break;
            case 'e': eps = static_cast<int>(std::stod(optarg)); break;
            case 'g': opt.gpus = std::stoi(optarg); break;
            case 'l': {
                const std::string l = optarg;
                if (l != "chain" && std::sscanf(optarg, "%dx%dx%d", &opt.lx, &opt.ly, &opt.lz) < 2) {
                    std::cerr << "Error: lattice must be 'chain' or LXxLY[xLZ]." << std::endl;
                    return 1;
                }
                break;
            }
            case 'k': opt.kernel = (std::string(optarg) == "free") ? 1 : 0; break;
            case 'o': opt.output = optarg; break;
            case 'c': opt.contexts_per_gpu = std::stoi(optarg); break;
            case 1000: plot = false; break;
            case 1001: opt.reuse_shift = true; break;
            case 1002: opt.resume = true; break;
            case 1003: opt.batch = std::max(1, std::min(4, std::atoi(optarg))); break;
            case 'h':
            default: print_usage(); return 0;
        }
    }
    if (calc_type != "exact" && calc_type != "mean") {
        std::cerr << "Error: calculation type must be 'exact' or 'mean'." << std::endl;
        return 1;
    }
    if (calc_type == "exact") {
        if (s >= r) {
            std::cerr << "Error: s must be smaller than r." << std::endl;
            return 1;
        }
        if (fixed_param != "J" && fixed_param != "U" && fixed_param != "u") {
            std::cerr << "Error: fixed parameter must be J, U or u." << std::endl;
            return 1;
        }
        Analysis::exact_parameters(m, n, J, U, mu, s, r, fixed_param, opt);
        if (plot && access("plot.py", R_OK) == 0 && std::system("python3 plot.py") != 0) {
            std::cerr << "Error when executing Python script." << std::endl;
            return 1;
        }
    } else {
        // the self-consistent mean-field mode (src/analysis.cpp:58-178) is outside the exact-diagonalisation hot path this
        // binary accelerates (SURVEY.md section 2 "out of scope"): distinct exit code so that scripts can tell it from a failure
        Analysis::mean_field_parameters(it, eps);
        return 2;
    }
    return 0;
}
