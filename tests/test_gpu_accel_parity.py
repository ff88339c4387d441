"""GPU: the *accelerated* solver (Chebyshev-filtered thick-restart Lanczos + Rayleigh-Ritz of H, csrc/lanczos.cu) --
the path every benchmarked number comes from -- pinned the way Spectra pins its own solvers
(external/spectra/test/SymEigs.cpp: residual of every returned pair), at the sizes of BASELINE.json's configs:

  * |H u - theta u|_inf <= 1e-9 and |u.u - 1| <= 1e-10 for all 20 vectors at m = n = 10 and 12 (C2, C3), for
    U values in every regime of the filter's `cut` rule, and for nev = 2 and nev = 20 at m = n = 14 (C5);
  * stored vs matrix-free H.v at m = n = 14 (the two kernels share no code path below the C ABI);
  * the lockstep-batched sweep path (bh_points, batch 4) against the single-point path and the golden points.
"""
import os

import numpy as np
import pytest

from parity_util import assert_out3

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))

RESID_TOL = 1e-9   # Spectra's own test bar (SymEigs.cpp), absolute, infinity norm
NORM_TOL = 1e-10


def check_pairs(pkg, ctx, pars, r, kernel):
    worst = 0.0
    for k in range(len(r["evals"])):
        u = r["vecs"][k]
        hu = ctx.hv(*pars, u, kernel=kernel, order=pkg.capi.LEX)
        res = np.abs(hu - r["evals"][k] * u).max()
        worst = max(worst, res)
        assert res <= RESID_TOL, (k, res)
        assert abs(np.dot(u, u) - 1.0) <= NORM_TOL, (k, np.dot(u, u))
    # pairwise orthogonality of the returned vectors (degenerate pairs included)
    V = np.asarray(r["vecs"])
    gram = V @ V.T
    assert np.abs(gram - np.eye(len(V))).max() <= 1e-9
    return worst


@pytest.mark.parametrize("m,U", [(10, 1.0), (10, 4.0), (10, 16.0), (10, 32.0), (12, 1.0), (12, 4.0), (12, 16.0), (12, 32.0)])
def test_accelerated_solver_residuals(pkg, ctx_factory, m, U):
    ctx = ctx_factory(m, m)
    pars = (1.0, U, 1.0)
    r = ctx.eigs(*pars, nev=20, kernel=pkg.capi.HV_MATRIX_FREE, order=pkg.capi.LEX, want_vectors=True)
    # the filtered path ran: one re-orthogonalisation serves a whole filter application (degree 8 by default)
    assert r["nreorth"] * 3 < r["nmatvec"], (r["nreorth"], r["nmatvec"])
    assert (np.diff(r["evals"]) >= -1e-9).all()
    check_pairs(pkg, ctx, pars, r, pkg.capi.HV_MATRIX_FREE)


def test_accelerated_solver_residuals_stored_kernel(pkg, ctx_factory):
    ctx = ctx_factory(10, 10)
    pars = (1.0, 8.0, 0.0)
    r = ctx.eigs(*pars, nev=20, kernel=pkg.capi.HV_STORED, order=pkg.capi.LEX, want_vectors=True)
    assert r["nreorth"] * 3 < r["nmatvec"]
    check_pairs(pkg, ctx, pars, r, pkg.capi.HV_STORED)


@pytest.fixture(scope="module")
def ctx14(pkg):
    c = pkg.Context(0).setup(14, 14)
    yield c
    c.close()


def test_c5_m14_stored_vs_matrix_free_hv(pkg, ctx14):
    D = ctx14.D
    assert D == 20058300
    x = np.random.default_rng(14).uniform(-0.5, 0.5, D)
    for pars in [(1.0, 4.0, 1.0), (0.7, 0.0, 2.0)]:
        a = ctx14.hv(*pars, x, kernel=pkg.capi.HV_MATRIX_FREE, order=pkg.capi.LEX)
        b = ctx14.hv(*pars, x, kernel=pkg.capi.HV_STORED, order=pkg.capi.LEX)
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    # symmetry of the matrix-free operator: <y, H x> = <H y, x>
    y = np.random.default_rng(15).uniform(-0.5, 0.5, D)
    hx = ctx14.hv(1.0, 4.0, 1.0, x, kernel=pkg.capi.HV_MATRIX_FREE, order=pkg.capi.LEX)
    hy = ctx14.hv(1.0, 4.0, 1.0, y, kernel=pkg.capi.HV_MATRIX_FREE, order=pkg.capi.LEX)
    assert abs(np.dot(y, hx) - np.dot(hy, x)) <= 1e-11 * abs(np.dot(y, hx))


@pytest.mark.parametrize("nev,ncv", [(2, 12), (20, 41)])
def test_c5_m14_residuals(pkg, ctx14, nev, ncv):
    pars = (1.0, 4.0, 1.0)
    r = ctx14.eigs(*pars, nev=nev, ncv=ncv, kernel=pkg.capi.HV_MATRIX_FREE, order=pkg.capi.LEX, want_vectors=True)
    assert r["nconv"] == nev
    check_pairs(pkg, ctx14, pars, r, pkg.capi.HV_MATRIX_FREE)
    # Rayleigh quotients through the *stored* kernel reproduce the eigenvalues (independent H.v code path)
    for k in range(min(nev, 3)):
        u = r["vecs"][k]
        hu = ctx14.hv(*pars, u, kernel=pkg.capi.HV_STORED, order=pkg.capi.LEX)
        assert abs(np.dot(u, hu) - r["evals"][k]) <= 1e-10 * abs(r["evals"][k])
    if nev == 20:
        # KA5 at full size: mu shifts the spectrum by -mu n and nothing else
        r2 = ctx14.eigs(1.0, 4.0, 0.0, nev=2, ncv=12, kernel=pkg.capi.HV_MATRIX_FREE, order=pkg.capi.LEX)
        assert np.abs((r2["evals"] - 14.0) - r["evals"][:2]).max() <= 1e-9 * abs(r["evals"][0])


M12_KEYS = sorted(k[len("point_"):-len("_evals")] for k in G.files
                  if k.startswith("point_12_12_") and k.endswith("_evals") and "rect" not in k)


def test_c3_lockstep_sweep_against_reference_points(pkg, ctx_factory):
    """The benchmarked path -- bh_points with 4 lockstep solves on the matrix-free kernel -- against every m = n = 12
    golden point of the compiled reference (U = 1, 4, 16, 32: every regime of the filter) and against the
    single-point path (bit-identical by construction)."""
    assert len(M12_KEYS) >= 4, M12_KEYS
    pars = np.array([[float(v) for v in k.split("_")[2:5]] for k in M12_KEYS])
    ctx = ctx_factory(12, 12)
    ctx.set_batch(4)
    out3, infos = ctx.points(pars[:, 0], pars[:, 1], pars[:, 2], kernel=pkg.capi.HV_MATRIX_FREE)
    ctx.set_batch(1)
    for i, key in enumerate(M12_KEYS):
        want = G[f"point_{key}_out5"][2:]
        single = ctx.point(*pars[i], kernel=pkg.capi.HV_MATRIX_FREE)
        assert (single["out3"] == out3[i]).all(), (key, single["out3"], out3[i])
        assert single["nmatvec"] == infos[i]["nmatvec"]
        ev = np.sort(G[f"point_{key}_evals"])
        scale = np.maximum(np.abs(ev), np.abs(ev[0]))
        assert np.all(np.abs(single["evals"] - ev) <= 1e-10 * scale), (key, single["evals"] - ev)
        assert_out3(out3[i], want, ev, single["evals"])   # gap ratio: tolerance propagated from the observed level differences
        rho = G[f"point_{key}_rho"]
        assert np.abs(single["rho"] - rho).max() <= 1e-10 * np.abs(rho).max()
