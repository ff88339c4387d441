"""GPU: the CUDA path against the committed golden fixtures (outputs of the real patched reference)."""
import os

import numpy as np
import pytest

from checksums import checksum_basis, checksum_csc
from parity_util import assert_out3

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


def parse_point_key(key):
    f = key.split("_")
    lat = f[5] if len(f) > 5 else "chain"
    return int(f[0]), int(f[1]), float(f[2]), float(f[3]), float(f[4]), lat


@pytest.mark.parametrize("key", sorted(POINTS))
def test_points(pkg, ctx_factory, key):
    """Every golden grid point of the compiled reference (chains m = 5..12 at U = 0.5..32, periodic rectangles 3x2, 4x2,
    3x3, 4x3 -- BASELINE.json config 4's geometry -- up to D = 31 824): 20 levels, rho, out3, both H.v kernels."""
    m, n, cJ, cU, cu, lat = parse_point_key(key)
    ctx = ctx_factory(m, n, nbr_of(pkg, lat, m))
    want = np.sort(G[f"point_{key}_evals"])
    if f"point_{key}_dense" in G.files:
        # Where dense diagonalisation of the same H is affordable the bar is the TRUTH.  It equals the reference's output at
        # every fixture except the 4 x 3 torus with n = 3, where the reference's single-vector Krylov solver misses copies of
        # two four-fold degenerate levels (tests/test_oracle_golden.py::test_reference_against_dense_truth pins that).
        want = G[f"point_{key}_dense"][:20]
    scale = np.maximum(np.abs(want), np.abs(want[0]))
    rho = G[f"point_{key}_rho"]
    for kernel in (pkg.capi.HV_STORED, pkg.capi.HV_MATRIX_FREE):
        got = ctx.point(cJ, cU, cu, kernel=kernel)
        assert np.all(np.abs(got["evals"] - want) <= 1e-10 * scale), (kernel, got["evals"] - want)
        assert np.abs(got["rho"] - rho).max() <= 1e-10 * np.abs(rho).max(), kernel
        assert_out3(got["out3"], G[f"point_{key}_out5"][2:], want, got["evals"])   # see parity_util (gap-ratio conditioning)


@pytest.mark.parametrize("m,n,lx,ly", [(6, 4, 3, 2), (8, 6, 4, 2), (12, 3, 4, 3)])
def test_rect_grid_through_points(pkg, ctx_factory, m, n, lx, ly):
    # a 3 x 3 corner of a sweep on a periodic rectangle through bh_points, vs the reference's loop body on the same neighbours
    out5 = G[f"grid_{m}_{n}_rect-{lx}-{ly}_out5"]
    evals = G[f"grid_{m}_{n}_rect-{lx}-{ly}_evals"]
    ctx = ctx_factory(m, n, pkg.capi.neighbours_rect(lx, ly))
    for kernel in (pkg.capi.HV_STORED, pkg.capi.HV_MATRIX_FREE):
        got, _ = ctx.points(np.ones(len(out5)), out5[:, 0], out5[:, 1], kernel=kernel)
        for i in range(len(out5)):
            e = ctx.point(1.0, out5[i, 0], out5[i, 1], kernel=kernel)["evals"]
            assert_out3(got[i], out5[i, 2:], evals[i], e)
