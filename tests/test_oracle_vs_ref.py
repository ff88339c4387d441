"""CPU, only where oracle/_ref exists (this container): the oracle restatement against the compiled patched
reference on shapes that are not in the committed fixtures."""
import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("m,n", [(4, 6), (6, 4), (7, 7)])
def test_basis_csc_hsum(m, n):
    t, b = O.basis(m, n)
    rt, rb, _ = R.basis(m, n)
    assert (bits(t) == bits(rt)).all() and (b == rb).all()
    t2, b2 = O.basis(m, n, O.REF_SCATTER)
    rt2, rb2, _ = R.basis(m, n, unpatched=True)
    assert (bits(t2) == bits(rt2)).all() and (b2 == rb2).all()
    jc = O.hopping_csc(m, O.chain(m), t, b)
    rj, _ = R.csc(m, n, "J")
    for a, c in zip(jc, rj):
        assert a.shape == c.shape and (a == c).all()
    dU, dN = O.diagonals(m, b)
    rh, _ = R.hsum(m, n, 0.7, 3.1, 0.9)
    for a, c in zip(O.hsum_csc(jc, dU, dN, 0.7, 3.1, 0.9), rh):
        assert a.shape == c.shape and (a == c).all()


def test_rect_lattice_point():
    m, n = 6, 4
    nbr = O.rect(3, 2)
    t, b = O.basis(m, n)
    jc = O.hopping_csc(m, nbr, t, b)
    rj, _ = R.csc(m, n, "J", "rect:3:2")
    for a, c in zip(jc, rj):
        assert (a == c).all()
    dU, dN = O.diagonals(m, b)
    r = O.point(m, t, b, jc, dU, dN, 1.0, 2.0, 0.5)
    rr, info = R.eigs(m, n, 1, 2, 0.5, lattice="rect:3:2")
    assert np.abs(r["evals"] - rr["evals"]).max() < 1e-10 * np.abs(rr["evals"]).max()
    assert np.allclose(r["out3"], rr["out5"][2:], rtol=1e-9)
    assert abs(r["nmatvec"] - info["nmatvec"]) <= 0.1 * info["nmatvec"]
