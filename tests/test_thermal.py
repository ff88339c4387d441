"""Finite-temperature branch of the sweep body (reference src/analysis.cpp:321-323, 456-494).  Dead code
in the reference (its temperature is the constant 0) but part of the path's source: pinned against the compiled reference
(oracle/_ref/ref_harness thermal -> tests/golden thermal_*)."""
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
    top = np.sort(np.linalg.eigvalsh(G[key + "_dm"]))[-20:]
    assert np.allclose(np.sort(w), top, rtol=1e-10, atol=1e-14)
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
    # u u^T is sign-invariant and, summed over a degenerate pair (equal weights), basis-invariant
    assert np.abs(dm - ref).max() <= 1e-10 * np.abs(ref).max()
    # the two scalars the reference's loop derives from it (:331-337)
    ev = np.linalg.eigvals(dm)
    cf = abs(ev[np.argmax(np.abs(ev))].real / np.trace(dm))
    K = (np.sum(dm * dm.T) - np.sum(np.diag(dm) ** 2)) / np.sum(dm * dm.T)
    assert np.allclose([cf, K], G[key + "_out2"], rtol=1e-9)
