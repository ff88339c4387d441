// small_dense_wrap.cpp -- TEST INFRASTRUCTURE ONLY: a C entry point around the library's host-side symmetric
// eigensolver (csrc/small_dense.cpp) so that the CPU test-suite can compare it with numpy.linalg.eigh.
#include <vector>
void bh_sym_eig(int n, std::vector<double>& a, std::vector<double>& evals, std::vector<double>& v);
extern "C" void wrap_sym_eig(int n, const double* a_colmajor, double* evals, double* vecs_colmajor)
{
    std::vector<double> a(a_colmajor, a_colmajor + (size_t)n * n), ev, v;
    bh_sym_eig(n, a, ev, v);
    for (int i = 0; i < n; ++i) evals[i] = ev[i];
    for (size_t i = 0; i < (size_t)n * n; ++i) vecs_colmajor[i] = v[i];
}
