"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.

Bars: basis / tags / ranks / sparsity pattern / values bit-exact;
H.v to 1e-13; eigenvalues |dE| <= 1e-10 * max(|E|, |E0|); observables 1e-10-level (1e-9 relative written below).
"""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

SHAPES = [(2, 3), (3, 2), (4, 4), (5, 5), (6, 6), (8, 8), (5, 9), (3, 15), (7, 3)]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("m,n", SHAPES)
def test_basis_three_orders_bit_exact(pkg, ctx_factory, m, n):
    ctx = ctx_factory(m, n)
    assert ctx.D == O.dimension(m, n)
    for order in (pkg.capi.LEX, pkg.capi.TAG_SORTED, pkg.capi.REF_SCATTER):
        tags, bas = ctx.basis(order)
        otags, obas = O.basis(m, n, order)
        assert (bits(tags) == bits(otags)).all()
        assert (bas == obas).all()


@pytest.mark.parametrize("m,n", [(4, 4), (6, 6), (8, 8), (5, 9)])
def test_rank_equals_search_tag(pkg, ctx_factory, m, n):
    ctx = ctx_factory(m, n)
    otags, obas = O.basis(m, n, O.TAG_SORTED)
    rng = np.random.default_rng(5)
    ks = rng.integers(0, len(otags), size=300)
    states, want = [], []
    for k in ks:
        s = obas[k].copy()
        occ = np.nonzero(s >= 1)[0]
        j = rng.choice(occ)
        i = rng.integers(0, m)
        s[i] += 1
        s[j] -= 1
        states.append(s)
        want.append(O.search_tag(otags, float(np.sum(s * np.log(np.array([2, 3, 5, 7, 11, 13, 17, 19, 23][:m], dtype=float))))))
    got = ctx.rank(np.array(states), pkg.capi.TAG_SORTED)
    # the oracle's search uses its own tag arithmetic; recompute the expected rank exactly from the basis too
    index = {tuple(r): i for i, r in enumerate(obas.astype(int))}
    exact = np.array([index[tuple(s.astype(int))] for s in states])
    assert (got == exact).all()
    assert (np.array(want) == exact).all()
    bad = np.array([[n + 1] + [0] * (m - 1), [-1] + [n + 1] + [0] * (m - 2)], dtype=float)
    assert (ctx.rank(bad, pkg.capi.TAG_SORTED) == -1).all()


LATTICES = [
    ("chain", 3, 2), ("chain", 4, 4), ("chain", 6, 6), ("chain", 8, 8), ("chain", 2, 5), ("chain", 5, 9),
    ("openchain", 5, 5), ("rect3x2", 6, 4), ("rect2x2", 4, 5), ("rect4x3", 12, 3),
]


def lattice(name, m):
    if name == "chain":
        return O.chain(m)
    if name == "openchain":
        return O.chain(m, closed=False)
    lx, ly = [int(v) for v in name[4:].split("x")]
    assert lx * ly == m
    return O.rect(lx, ly)


@pytest.mark.parametrize("name,m,n", LATTICES)
def test_csc_pattern_and_values_bit_exact(pkg, ctx_factory, name, m, n):
    nbr = lattice(name, m)
    ctx = ctx_factory(m, n, nbr)
    otags, obas = O.basis(m, n, O.TAG_SORTED)
    jc = O.hopping_csc(m, nbr, otags, obas, 1.0)
    dU, dN = O.diagonals(m, obas)
    got = ctx.term_csc(pkg.capi.TERM_J, 1.0, pkg.capi.TAG_SORTED)
    for a, b in zip(got, jc):
        assert a.shape == b.shape and (a == b).all()
    # coefficient semantics of fixed_bosons_hamiltonian(.., J = 0.37, 0, 0)
    jc2 = O.hopping_csc(m, nbr, otags, obas, 0.37)
    got2 = ctx.term_csc(pkg.capi.TERM_J, 0.37, pkg.capi.TAG_SORTED)
    assert (got2[2] == jc2[2]).all()
    gu = ctx.term_csc(pkg.capi.TERM_U, 1.0, pkg.capi.TAG_SORTED)
    assert (gu[0] == np.arange(len(dU) + 1)).all() and (gu[1] == np.arange(len(dU))).all() and (gu[2] == dU).all()
    gm = ctx.term_csc(pkg.capi.TERM_MU, 1.0, pkg.capi.TAG_SORTED)
    assert (gm[2] == dN).all()
    for (cJ, cU, cmu) in [(1.0, 4.0, 1.0), (0.5, 2.5, 0.0), (1.0, 1.0 / 3.0, 7.25)]:
        want = O.hsum_csc(jc, dU, dN, cJ, cU, cmu)
        goth = ctx.hamiltonian_csc(cJ, cU, cmu, pkg.capi.TAG_SORTED)
        for a, b in zip(goth, want):
            assert a.shape == b.shape and (a == b).all()


def test_known_answers_pattern(pkg, ctx_factory):
    # KA1/KA2: D and nnz(JH) = D z m n / (m + n - 1), symmetric, no diagonal
    for m, n, nnz in [(8, 8, 54912), (10, 10, 972400)]:
        ctx = ctx_factory(m, n)
        assert ctx.D == {8: 6435, 10: 92378}[m]
        assert ctx.term_nnz(pkg.capi.TERM_J) == nnz
        assert ctx.hamiltonian_nnz() == nnz + ctx.D


@pytest.mark.parametrize("name,m,n", [("chain", 6, 6), ("chain", 8, 8), ("rect3x2", 6, 5), ("rect2x2", 4, 6), ("chain", 2, 7)])
def test_hv_both_kernels(pkg, ctx_factory, name, m, n):
    nbr = lattice(name, m)
    ctx = ctx_factory(m, n, nbr)
    otags, obas = O.basis(m, n, O.TAG_SORTED)
    jc = O.hopping_csc(m, nbr, otags, obas, 1.0)
    dU, dN = O.diagonals(m, obas)
    x = O.lcg_vector(len(otags))
    for (cJ, cU, cmu) in [(1.0, 4.0, 1.0), (0.3, 0.0, 2.0)]:
        h = O.hsum_csc(jc, dU, dN, cJ, cU, cmu)
        want = O.spmv(h, x)
        scale = np.abs(want).max()
        for kernel in (pkg.capi.HV_STORED, pkg.capi.HV_MATRIX_FREE):
            got = ctx.hv(cJ, cU, cmu, x, kernel=kernel, order=pkg.capi.TAG_SORTED)
            assert np.abs(got - want).max() <= 1e-13 * scale, (kernel, np.abs(got - want).max())
        # LEX order: same operator in the device's own ordering
        _, lbas = O.basis(m, n, O.LEX)
        perm = ctx.rank(lbas, pkg.capi.TAG_SORTED)  # lex state -> tag position
        got_lex = ctx.hv(cJ, cU, cmu, x[perm], kernel=pkg.capi.HV_STORED, order=pkg.capi.LEX)
        assert np.abs(got_lex - want[perm]).max() <= 1e-13 * scale


def eig_close(got, want):
    scale = np.maximum(np.abs(want), np.abs(want[0]))
    return np.all(np.abs(np.sort(got) - np.sort(want)) <= 1e-10 * scale)


@pytest.mark.parametrize("m,n,pars", [(5, 5, (1.0, 4.0, 1.0)), (6, 6, (1.0, 4.0, 1.0)), (6, 6, (1.0, 0.5, 0.0)),
                                      (8, 8, (1.0, 4.0, 1.0)), (8, 8, (1.0, 10.0, 0.0)), (7, 6, (1.0, 2.0, 3.0))])
def test_point_eigenvalues_and_observables(pkg, ctx_factory, m, n, pars):
    ctx = ctx_factory(m, n)
    otags, obas = O.basis(m, n, O.TAG_SORTED)
    jc = O.hopping_csc(m, O.chain(m), otags, obas, 1.0)
    dU, dN = O.diagonals(m, obas)
    ref = O.point(m, otags, obas, jc, dU, dN, *pars)
    for kernel in (pkg.capi.HV_STORED, pkg.capi.HV_MATRIX_FREE):
        got = ctx.point(*pars, kernel=kernel)
        assert eig_close(got["evals"], ref["evals"]), (got["evals"] - ref["evals"])
        assert np.allclose(got["rho"], ref["rho"], rtol=0, atol=1e-10 * np.abs(ref["rho"]).max())
        assert np.allclose(got["out3"], ref["out3"], rtol=1e-9, atol=1e-12), (got["out3"], ref["out3"])


def test_known_answers_spectrum(pkg, ctx_factory):
    # KA3: U = 0 -> E0 = -4 J n, next level E0 + 4J(1 - cos 2pi/m) twice
    m = n = 8
    ctx = ctx_factory(m, n)
    r = ctx.eigs(1.0, 0.0, 0.0, nev=20)
    e = np.sort(r["evals"])
    assert abs(e[0] + 32.0) < 1e-9
    gap = 4 * (1 - np.cos(2 * np.pi / m))
    assert abs(e[1] - (-32 + gap)) < 1e-9 and abs(e[2] - (-32 + gap)) < 1e-9
    # KA5: mu only shifts the spectrum by -mu n
    a = ctx.eigs(1.0, 4.0, 0.0)["evals"]
    b = ctx.eigs(1.0, 4.0, 2.5)["evals"]
    assert np.allclose(b, a - 2.5 * n, rtol=0, atol=1e-9)


def test_eigenvectors_residual(pkg, ctx_factory):
    # the Spectra-style pin: |H u - theta u| <= 1e-9 (external/spectra/test/SymEigs.cpp)
    m = n = 6
    ctx = ctx_factory(m, n)
    r = ctx.eigs(1.0, 4.0, 1.0, nev=20, want_vectors=True)
    for k in range(20):
        u = r["vecs"][k]
        hu = ctx.hv(1.0, 4.0, 1.0, u, order=pkg.capi.TAG_SORTED)
        assert np.abs(hu - r["evals"][k] * u).max() <= 1e-9
        assert abs(np.dot(u, u) - 1) < 1e-10


def test_error_behaviour(pkg, ctx_factory):
    # D(4,4) = 35 < ncv = 41: Spectra throws "ncv must satisfy nev + 2 <= ncv <= n" 
    ctx = ctx_factory(4, 4)
    with pytest.raises(pkg.BhError) as ei:
        ctx.point(1.0, 4.0, 1.0)
    assert ei.value.code == pkg.capi.ERR_ARG and "ncv must satisfy" in str(ei.value)
    with pytest.raises(pkg.BhError):
        pkg.Context(0).setup(17, 3)
