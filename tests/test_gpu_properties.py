"""GPU: BASELINE.json's full sizes through size-independent properties (the oracle cannot run these in seconds):
KA1/KA2 counts, symmetry <y, Hx> = <x, Hy>, linearity, the mu-shift identity (KA5), stored == matrix-free,
20 tr(rho) = n and rho_ii = n/m/20 (KA6), scale invariance (KA8), golden eigenvalues at m = n = 12."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def test_c3_counts_and_hv_properties(pkg, ctx_factory):
    m = n = 12
    c = pkg.capi
    ctx = ctx_factory(m, n)
    D = ctx.D
    assert D == 1352078 and ctx.term_nnz(c.TERM_J) == 16930368 and ctx.hamiltonian_nnz() == 18282446
    assert ctx.hv_algorithmic_bytes(c.HV_STORED) == 246430916 and ctx.hv_algorithmic_bytes(c.HV_MATRIX_FREE) == 16 * D
    rng = np.random.default_rng(7)
    x, y = rng.uniform(-0.5, 0.5, D), rng.uniform(-0.5, 0.5, D)
    for kernel in (c.HV_STORED, c.HV_MATRIX_FREE):
        hx = ctx.hv(1.0, 4.0, 1.0, x, kernel=kernel, order=c.LEX)
        hy = ctx.hv(1.0, 4.0, 1.0, y, kernel=kernel, order=c.LEX)
        assert abs(np.dot(y, hx) - np.dot(x, hy)) <= 1e-11 * abs(np.dot(y, hx))           # symmetric
        hxy = ctx.hv(1.0, 4.0, 1.0, 2.0 * x - 3.0 * y, kernel=kernel, order=c.LEX)
        assert np.abs(hxy - (2.0 * hx - 3.0 * hy)).max() <= 1e-12 * np.abs(hxy).max()     # linear
        h0 = ctx.hv(1.0, 4.0, 0.0, x, kernel=kernel, order=c.LEX)
        assert np.abs(hx - (h0 - 1.0 * n * x)).max() <= 1e-12 * np.abs(hx).max()           # KA5: -mu n I
    a = ctx.hv(0.7, 2.5, 0.3, x, kernel=c.HV_STORED, order=c.LEX)
    b = ctx.hv(0.7, 2.5, 0.3, x, kernel=c.HV_MATRIX_FREE, order=c.LEX)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()
    # tag-sorted order is the same operator conjugated by the permutation
    t = ctx.hv(0.7, 2.5, 0.3, x, kernel=c.HV_STORED, order=c.TAG_SORTED)
    t2 = ctx.hv(0.7, 2.5, 0.3, x, kernel=c.HV_MATRIX_FREE, order=c.TAG_SORTED)
    assert np.abs(t - t2).max() <= 1e-13 * np.abs(t).max()


def test_c3_point_against_reference_and_invariances(pkg, ctx_factory):
    m = n = 12
    ctx = ctx_factory(m, n)
    r = ctx.point(1.0, 4.0, 1.0)
    e = r["evals"]
    key = "point_12_12_1_4_1_evals"
    if key in G.files:
        want = np.sort(G[key])
        assert np.all(np.abs(e - want) <= 1e-10 * np.maximum(np.abs(want), abs(want[0])))
        assert np.allclose(r["out3"], G["point_12_12_1_4_1_out5"][2:], rtol=1e-9, atol=1e-12)
    # survey-time values of the patched reference
    assert abs(e[0] - 61.8911362495153) < 1e-8 and abs(e[1] - 64.6269331073841) < 1e-8 and abs(e[2] - 64.6269331073841) < 1e-8
    assert abs(e[3] - 65.6994013934106) < 1e-8
    rho = r["rho"]
    assert abs(20 * np.trace(rho) - n) < 1e-9 and np.abs(rho - rho.T).max() < 1e-14       # KA6
    assert np.abs(20 * np.diag(rho) - n / m).max() < 1e-8
    r2 = ctx.point(2.0, 8.0, 0.0)                                                          # KA8 + KA5
    assert np.allclose(r2["out3"], r["out3"], rtol=1e-8)
    assert np.allclose(r2["evals"] / 2.0, e + 1.0 * n, rtol=0, atol=1e-8)


def test_c4_rect_lattice(pkg, ctx_factory):
    c = pkg.capi
    ctx = ctx_factory(12, 12, c.neighbours_rect(4, 3))
    assert ctx.term_nnz(c.TERM_J) == 33860736 and ctx.hv_algorithmic_bytes(c.HV_STORED) == 449595332
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.5, 0.5, ctx.D)
    a = ctx.hv(1.0, 4.0, 1.0, x, kernel=c.HV_STORED, order=c.LEX)
    b = ctx.hv(1.0, 4.0, 1.0, x, kernel=c.HV_MATRIX_FREE, order=c.LEX)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()
    e = ctx.eigs(1.0, 0.0, 0.0, nev=2, ncv=12)["evals"]
    assert abs(e[0] + 2 * 4 * 12) < 1e-8                                                   # U = 0: E0 = -2 J z n


def test_c5_matrix_free_m14(pkg, ctx_factory):
    c = pkg.capi
    ctx = ctx_factory(14, 14)
    D = ctx.D
    assert D == 20058300 and ctx.hv_algorithmic_bytes(c.HV_MATRIX_FREE) == 320932800
    rng = np.random.default_rng(11)
    x, y = rng.uniform(-0.5, 0.5, D), rng.uniform(-0.5, 0.5, D)
    hx = ctx.hv(1.0, 4.0, 1.0, x, kernel=c.HV_MATRIX_FREE, order=c.LEX)
    hy = ctx.hv(1.0, 4.0, 1.0, y, kernel=c.HV_MATRIX_FREE, order=c.LEX)
    assert abs(np.dot(y, hx) - np.dot(x, hy)) <= 1e-11 * abs(np.dot(y, hx))
    e = ctx.eigs(1.0, 0.0, 0.0, nev=2, ncv=12, kernel=c.HV_MATRIX_FREE)["evals"]         # KA3 at the largest size
    assert abs(e[0] + 56.0) < 1e-8 and abs(e[1] - (-56 + 4 * (1 - np.cos(2 * np.pi / 14)))) < 1e-8
