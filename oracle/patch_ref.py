#!/usr/bin/env python3
"""Generate the *patched* reference sources in a scratch directory (default /tmp/bh_ref_src).

TEST INFRASTRUCTURE ONLY.  Nothing here is shipped or imported by the product.

The reference's exact-diagonalisation path has undefined behaviour / races
(defects D1..D6).  The oracle policy is:
keep every deterministic quirk, patch only what is UB, racy or non-functional.
This script reads the sources where they lie (/root/reference/src), applies the
six small patches P1..P6 by exact-string replacement (it fails loudly if the
reference text is not what was surveyed) and writes the result to the scratch
directory given as argv[1].  Reference sources are never copied into the repo:
only the binaries built from them land in oracle/_ref/ (git-ignored).

P1  hamiltonian.cpp:180, analysis.cpp:581  calculate_tag(state, primes, i) -> (.., 0)
P2  hamiltonian.cpp:115-122                sort_basis applies the permutation as a gather
P3  hamiltonian.cpp:177-183                source site is neighbours[i][j], not j
P4  hamiltonian.cpp:130                    search_tag tolerance 1e-3 -> 1e-12
P5  analysis.cpp:314                       per-iteration eigenvectors (no shared write)
P6  operator.cpp:26                        final sorting SmallestReal (col 0 = ground state)
"""
import os
import sys

REF = os.environ.get("BH_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = sys.argv[1] if len(sys.argv) > 1 else "/tmp/bh_ref_src"


def sub(text, old, new, count=1, what=""):
    n = text.count(old)
    if n != count:
        raise SystemExit(f"patch_ref: expected {count} occurrence(s) of {what or old!r}, found {n}")
    return text.replace(old, new)


def main():
    os.makedirs(OUT, exist_ok=True)
    src = lambda f: open(os.path.join(REF, "src", f)).read()

    # ---- hamiltonian.cpp: P1 (one site), P2, P3, P4 ----
    h = src("hamiltonian.cpp")
    orig_h = h
    h = sub(h, "calculate_tag(state, primes, i)", "calculate_tag(state, primes, 0)", what="P1/hamiltonian")
    # P2: replace the cycle-walk by an out-of-place gather new[i] = old[indices[i]]
    start = h.index("    for (int i = 0; i < static_cast<int>(indices.size()); ++i) {\n        while (indices[i] != i) {")
    end = h.index("/* Gives the index of the wanted tag x")
    gather = (
        "    { /* P2: gather */\n"
        "        Eigen::VectorXd t2(tags.size());\n"
        "        Eigen::MatrixXd b2(basis.rows(), basis.cols());\n"
        "        for (int i = 0; i < static_cast<int>(indices.size()); ++i) {\n"
        "            t2[i] = tags[indices[i]];\n"
        "            b2.col(i) = basis.col(indices[i]);\n"
        "        }\n"
        "        tags.swap(t2);\n"
        "        basis.swap(b2);\n"
        "    }\n"
        "}\n\n"
    )
    h = h[:start] + gather + h[end:]
    # P3: the source site of the hop is neighbours[i][j]
    h = sub(h, "if (basis.coeff(i, k) >= 0 && basis.coeff(j, k) >= 1) {",
            "const int src_site = neighbours[i][j];\n                if (basis.coeff(i, k) >= 0 && basis.coeff(src_site, k) >= 1) {",
            what="P3/guard")
    h = sub(h, "state[j] -= 1;\n                    double x = calculate_tag",
            "state[src_site] -= 1;\n                    double x = calculate_tag", what="P3/decrement")
    h = sub(h, "sqrt((basis.coeff(i, k) + 1) * basis.coeff(j, k))",
            "sqrt((basis.coeff(i, k) + 1) * basis.coeff(src_site, k))", what="P3/amplitude")
    # P4
    h = sub(h, "fabs(tags[m] - x) > 1e-3", "fabs(tags[m] - x) > 1e-12", what="P4")
    open(os.path.join(OUT, "hamiltonian.cpp"), "w").write(h)
    # the unpatched file is also needed: BH::fixed_set_basis is well defined as is
    # and its (scatter) order is a parity target of its own
    open(os.path.join(OUT, "hamiltonian_unpatched.cpp"), "w").write(orig_h)

    # ---- analysis.cpp: P1 (other site), P5 ----
    a = src("analysis.cpp")
    a = sub(a, "BH::calculate_tag(state, primes, i)", "BH::calculate_tag(state, primes, 0)", what="P1/analysis")
    a = sub(a, "Eigen::VectorXcd eigenvalues = Op::IRLM_eigen(H, nb_eigen, eigenvectors);",
            "Eigen::MatrixXcd eigenvectors; /* P5 */\n"
            "                Eigen::VectorXcd eigenvalues = Op::IRLM_eigen(H, nb_eigen, eigenvectors);",
            what="P5")
    open(os.path.join(OUT, "analysis.cpp"), "w").write(a)

    # ---- operator.cpp: P6 ----
    o = src("operator.cpp")
    o = sub(o, "eigs.compute(Spectra::SortRule::SmallestReal)",
            "eigs.compute(Spectra::SortRule::SmallestReal, 1000, 1e-10, Spectra::SortRule::SmallestReal)",
            what="P6")
    open(os.path.join(OUT, "operator.cpp"), "w").write(o)

    # ---- untouched translation units ----
    for f in ("neighbours.cpp", "resource.cpp", "main.cpp"):
        open(os.path.join(OUT, f), "w").write(src(f))
    print(f"patch_ref: wrote patched sources to {OUT}")


if __name__ == "__main__":
    sys.exit(main())
