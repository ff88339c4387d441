// hamiltonian.hpp -- drop-in for the externally linkable part of the reference's namespace BH
// (include/hamiltonian.hpp:33-100), same names / argument meaning / value semantics, Eigen types at the
// boundary, all heavy work behind the C ABI (K1, K2).  Needs Eigen headers (not vendored in this repo).
#pragma once

#include <Eigen/Dense>
#include <Eigen/SparseCore>
#include <utility>
#include <vector>

#include "../../include/bh_b200.h"

namespace BH {

// tag of column k (src/hamiltonian.cpp:91-97) / of every column (:100-106); small host helpers
double calculate_tag(const Eigen::MatrixXd& basis, const std::vector<int>& primes, int k);
Eigen::VectorXd calculate_tags(const Eigen::MatrixXd& basis, const std::vector<int>& primes);
// index of tag x in an ascending tag vector (:126-140, with the tolerance fix P4 of SURVEY.md)
int search_tag(const Eigen::VectorXd& tags, double x);

// (tags, basis) of the fixed-N Fock space (:143-149).  Order: ascending tag by default (what sort_basis
// intends); set_basis_order(BH_ORDER_REF_SCATTER) reproduces the order the unmodified reference returns.
std::pair<Eigen::VectorXd, Eigen::MatrixXd> fixed_set_basis(int m, int n);
std::pair<Eigen::VectorXd, Eigen::MatrixXd> max_set_basis(int m, int n);  // :152-166

// exactly one of the hopping / interaction / chemical terms, chosen like the reference (:238-256).  `basis`
// fixes the state order of the result; it must be a basis returned by fixed_set_basis (any of the three orders).
Eigen::SparseMatrix<double> fixed_bosons_hamiltonian(const std::vector<std::vector<int>>& neighbours,
                                                     const Eigen::MatrixXd& basis, const Eigen::VectorXd& tags, int m,
                                                     int n, double J, double U, double mu);
Eigen::SparseMatrix<double> max_bosons_hamiltonian(const std::vector<std::vector<int>>& neighbours, int m, int n_min,
                                                   int n_max, double J, double U, double mu);  // :260-288

// single-site mean-field Hamiltonian (:291-310): not on the accelerated path, a small host helper kept so that the
// reference's own analysis.cpp links against this layer unchanged (INTEGRATION.md option B)
void h_MF(double psi, int p, double mu, double J, int q, Eigen::MatrixXd& h);

// ---- shim controls (not in the reference) ----
void set_basis_order(int bh_order);
void set_device(int device);

}  // namespace BH
