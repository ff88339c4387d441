// operator.hpp -- drop-in for Op::IRLM_eigen (reference include/operator.hpp:34, src/operator.cpp:22-33).
// The never-called helpers of the reference (FOLM_diag / FOLM_eigen / exact_eigen, src/operator.cpp:39-101)
// are not reproduced.  Needs Eigen headers.
#pragma once

#include <Eigen/Dense>
#include <Eigen/SparseCore>

namespace Op {

// nb_eigen smallest eigenvalues (ascending, imaginary parts zero) of the real symmetric sparse matrix O and
// their eigenvectors (columns).  O is copied to the GPU (bh_load_matrix), solved by the device Lanczos with
// Spectra's parameters (ncv = 2 nb_eigen + 1, tol 1e-10, maxit 1000) and the vectors are copied back.
// Throws std::runtime_error("Eigenvalue computation failed.") like the reference, std::invalid_argument when
// Spectra's constructor would (nev + 2 <= ncv <= n violated).
Eigen::VectorXcd IRLM_eigen(Eigen::SparseMatrix<double> O, int nb_eigen, Eigen::MatrixXcd& eigenvectors);

void set_device(int device);

}  // namespace Op
