"""Finite-temperature branch of the sweep body (reference src/analysis.cpp:321-323, 456-494).  Dead code
in the reference (its temperature is the constant 0) but part of the path's source: pinned against the compiled reference
(oracle/_ref/ref_harness thermal -> tests/golden thermal_*).

What can be pinned: the reference builds rho = V diag(w) V^T from the eigenvectors of Spectra's GENERAL (non-symmetric)
solver, which are normalised but NOT orthogonal inside a degenerate eigenspace (momentum +k / -k pairs).  Its matrix is
therefore not a function of the spectrum and the eigenspaces alone: a degenerate pair with weight w contributes
w (a a^T + b b^T) with a.b != 0, whose two non-zero eigenvalues are w (1 +- a.b) instead of w, w.  The Boltzmann weights,
the trace over every eigenspace and everything built from non-degenerate levels are well defined and are what is compared;
the product's matrix uses orthonormal eigenvectors, i.e. it is sum_g w_g P_g with P_g the eigenspace projectors."""


def groups(evals, tol=1e-9):
    """index groups of (numerically) equal eigenvalues, ascending"""
    order = np.argsort(evals)
    out, cur = [], [order[0]]
    for a, b in zip(order[:-1], order[1:]):
        if abs(evals[b] - evals[a]) <= tol * max(1.0, abs(evals[a])):
            cur.append(b)
        else:
            out.append(cur)
            cur = [b]
    out.append(cur)
    return out
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
CASES = [(5, 5, 0.5), (5, 5, 3.0), (6, 4, 1.0)]


@pytest.mark.parametrize("m,n,T", CASES)
def test_thermal_weights_against_reference(pkg, m, n, T):
    # host-only entry point: the non-zero eigenvalues of the reference's density matrix ARE its Boltzmann weights
    key = f"thermal_{m}_{n}_{T:g}"
    w = pkg.capi.thermal_weights(G[key + "_evals"], T)
    assert abs(w.sum() - 1.0) <= 1e-14
    ev = G[key + "_evals"]
    top = np.sort(np.linalg.eigvalsh(G[key + "_dm"]))[-20:][::-1]   # descending = ascending energy
    assert abs(top.sum() - 1.0) <= 1e-12
    # the non-degenerate ground level is an exact eigenvalue of the reference's matrix; the others are pinned through the
    # eigenspace traces in the GPU test below (here: the split of the degenerate pairs stays within a few per cent)
    assert abs(top[0] - w[np.argmin(ev)]) <= 1e-10 * top[0]
    assert np.allclose(np.sort(top), np.sort(w), rtol=0.05)
    with pytest.raises(pkg.BhError):
        pkg.capi.thermal_weights(G[key + "_evals"], 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,T", CASES)
def test_density_matrix_against_reference(pkg, ctx_factory, m, n, T):
    key = f"thermal_{m}_{n}_{T:g}"
    ctx = ctx_factory(m, n)
    r = ctx.eigs(1.0, 4.0, 1.0, nev=20, kernel=pkg.capi.HV_STORED, order=pkg.capi.TAG_SORTED, want_vectors=True)
    want = np.sort(G[key + "_evals"])
    assert np.all(np.abs(r["evals"] - want) <= 1e-10 * np.maximum(np.abs(want), abs(want[0])))
    dm = ctx.density_matrix(r["evals"], r["vecs"], T)
    ref = G[key + "_dm"]
    w = pkg.capi.thermal_weights(r["evals"], T)
    V = np.asarray(r["vecs"])                      # rows = orthonormal eigenvectors
    assert np.abs(dm - dm.T).max() <= 1e-15 and abs(np.trace(dm) - 1.0) <= 1e-12
    assert np.abs(dm - (V.T * w) @ V).max() <= 1e-14     # the kernel computes sum_k w_k u_k u_k^T
    for g in groups(r["evals"]):
        if 19 in g:
            continue   # the 20th level may be one half of a degenerate pair cut by nev: its eigenvector is then not unique
        P = V[g].T @ V[g]                                 # eigenspace projector
        # trace of the reference's matrix over every eigenspace = the group's total weight (see the module docstring)
        assert abs(np.trace(P @ ref) - w[g].sum()) <= 1e-9 * w[g].sum(), g
        assert abs(np.trace(P @ dm) - w[g].sum()) <= 1e-12
        if len(g) == 1:   # non-degenerate level: the matrices agree on it completely
            u = V[g[0]]
            assert np.abs(ref @ u - dm @ u).max() <= 1e-9 * w[g[0]]
