"""GPU: the CUDA path against the committed golden fixtures (outputs of the real patched reference)."""
import os

import numpy as np
import pytest

from checksums import checksum_basis, checksum_csc

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("m,n", [(3, 2), (4, 4), (5, 3)])
def test_basis(pkg, ctx_factory, m, n):
    ctx = ctx_factory(m, n)
    for order, name in ((pkg.capi.TAG_SORTED, "sorted"), (pkg.capi.REF_SCATTER, "scatter")):
        t, b = ctx.basis(order)
        assert (bits(t) == bits(G[f"basis_{m}_{n}_{name}_tags"])).all() and (b == G[f"basis_{m}_{n}_{name}_states"]).all()


@pytest.mark.parametrize("m,n", [(6, 6), (8, 8), (10, 10)])
def test_basis_checksums(pkg, ctx_factory, m, n):
    ctx = ctx_factory(m, n)
    for order, name in ((pkg.capi.TAG_SORTED, "sorted"), (pkg.capi.REF_SCATTER, "scatter")):
        t, b = ctx.basis(order)
        assert (checksum_basis(t, b) == G[f"basissum_{m}_{n}_{name}"]).all()


def nbr_of(pkg, lat, m):
    if lat == "chain":
        return pkg.capi.neighbours_chain(m)
    lx, ly = [int(v) for v in lat.split("-")[1:]]
    return pkg.capi.neighbours_rect(lx, ly)


@pytest.mark.parametrize("m,n,lat", [(3, 2, "chain"), (4, 4, "chain"), (5, 5, "chain"), (2, 5, "chain"), (6, 3, "rect-3-2"),
                                     (4, 3, "rect-2-2")])
def test_csc(pkg, ctx_factory, m, n, lat):
    ctx = ctx_factory(m, n, nbr_of(pkg, lat, m))
    tag = f"{m}_{n}_{lat}"
    for term, name in ((pkg.capi.TERM_J, "J"), (pkg.capi.TERM_U, "U"), (pkg.capi.TERM_MU, "u")):
        got = ctx.term_csc(term, 1.0)
        for a, nm in zip(got, ("outer", "inner", "val")):
            assert (a == G[f"csc_{name}_{tag}_{nm}"]).all()
    got = ctx.hamiltonian_csc(1.0, 4.0, 1.0)
    for a, nm in zip(got, ("outer", "inner", "val")):
        assert (a == G[f"hsum_{tag}_{nm}"]).all()


@pytest.mark.parametrize("m,n", [(8, 8), (10, 10)])
def test_csc_checksum(pkg, ctx_factory, m, n):
    ctx = ctx_factory(m, n)
    o, i, v = ctx.term_csc(pkg.capi.TERM_J, 1.0)
    assert (checksum_csc(o, i, v) == G[f"cscsum_J_{m}_{n}"]).all()


@pytest.mark.parametrize("m,n", [(6, 6), (8, 8)])
def test_hv_vs_spectra_matop(pkg, ctx_factory, m, n):
    ctx = ctx_factory(m, n)
    x, want = G[f"hv_{m}_{n}_x"], G[f"hv_{m}_{n}_y"]
    for kernel in (pkg.capi.HV_STORED, pkg.capi.HV_MATRIX_FREE):
        y = ctx.hv(1.0, 4.0, 1.0, x, kernel=kernel)
        assert np.abs(y - want).max() <= 1e-13 * np.abs(want).max()


POINTS = [k[len("point_"):-len("_evals")] for k in G.files if k.startswith("point_") and k.endswith("_evals")]


@pytest.mark.parametrize("key", sorted(POINTS))
def test_points(pkg, ctx_factory, key):
    m, n, cJ, cU, cu = key.split("_")
    m, n, cJ, cU, cu = int(m), int(n), float(cJ), float(cU), float(cu)
    ctx = ctx_factory(m, n)
    got = ctx.point(cJ, cU, cu)
    want = G[f"point_{key}_evals"]
    scale = np.maximum(np.abs(want), np.abs(want[0]))
    assert np.all(np.abs(got["evals"] - np.sort(want)) <= 1e-10 * scale), got["evals"] - want
    rho = G[f"point_{key}_rho"]
    assert np.abs(got["rho"] - rho).max() <= 1e-10 * np.abs(rho).max()
    assert np.allclose(got["out3"], G[f"point_{key}_out5"][2:], rtol=1e-9, atol=1e-12)
