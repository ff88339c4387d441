"""GPU: the many-point small-system solver (csrc/small.cu: one CTA per grid point, a whole restart cycle per launch) -- the
path bh_points takes for BASELINE.json's configs 1 and 2 -- against the compiled reference's phase.txt and against the
single-point path."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def read_phase(name):
    lines = [l for l in open(os.path.join(GOLD, name)).read().split("\n") if l]
    return np.array([[float(v) for v in l.split()] for l in lines[1:]])


@pytest.mark.parametrize("name,m", [("phase_m8_C1.txt", 8), ("phase_m10_fJ.txt", 10), ("phase_m6_fJ.txt", 6), ("phase_m5_fJ.txt", 5)])
def test_small_solver_reproduces_reference_phase_txt(pkg, ctx_factory, name, m):
    """C1 in full (121 points, filtered solver), a corner of C2 (quick stage 1), and two tiny systems (plain mode)."""
    ref = read_phase(name)
    ctx = ctx_factory(m, m)
    out3, infos = ctx.points(np.ones(len(ref)), ref[:, 0], ref[:, 1], kernel=pkg.capi.HV_MATRIX_FREE)
    assert np.allclose(out3, ref[:, 2:], rtol=6e-6, atol=1e-9)   # the file holds 6 significant digits
    assert all(i["nmatvec"] > 0 for i in infos)


@pytest.mark.parametrize("m,n", [(8, 8), (9, 7), (10, 10)])
def test_small_solver_equals_single_points(pkg, ctx_factory, m, n):
    cU = np.array([1.0, 4.0, 2.5, 9.0, 6.0, 3.0, 12.0, 32.0, 0.5, 16.0, 7.0, 21.0])
    cmu = np.array([0.0, 1.0, 2.0, 0.5, 0.0, 3.0, 1.0, 0.0, 2.0, 5.0, 1.5, 0.25])
    cJ = np.ones_like(cU)
    ctx = ctx_factory(m, n)
    out3, infos = ctx.points(cJ, cU, cmu, kernel=pkg.capi.HV_MATRIX_FREE)
    for i in range(len(cU)):
        one = ctx.point(cJ[i], cU[i], cmu[i], kernel=pkg.capi.HV_MATRIX_FREE)
        # condensate fraction / coherence: functions of the ground vector; gap ratio: of level spacings (1e-12-level
        # differences of the two solvers' eigenvalues over spacings of 1e-2..1)
        assert np.allclose(out3[i, 1:], one["out3"][1:], rtol=1e-9, atol=1e-12), (i, out3[i], one["out3"])
        assert abs(out3[i, 0] - one["out3"][0]) <= 1e-8 * max(abs(one["out3"][0]), 1e-3), (i, out3[i], one["out3"])
        # same algorithm, same start vector: the H.v counts agree up to the restart-trajectory sensitivity
        assert abs(infos[i]["nmatvec"] - one["nmatvec"]) <= 0.25 * one["nmatvec"]


def test_small_solver_hands_breakdowns_to_the_single_point_path(pkg, ctx_factory):
    # J = 0 (diagonal H, invariant Krylov spaces) inside a list: that point falls back, the others are unaffected
    m = n = 6
    ctx = ctx_factory(m, n)
    cJ = np.array([1.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    cU = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0])
    cmu = np.full(9, 0.5)
    out3, _ = ctx.points(cJ, cU, cmu, kernel=pkg.capi.HV_MATRIX_FREE)
    for i in range(9):
        one = ctx.point(cJ[i], cU[i], cmu[i], kernel=pkg.capi.HV_MATRIX_FREE)
        if cJ[i] == 0.0:
            assert np.all(np.isfinite(out3[i]))   # multiplicities are not well defined at J = 0 (see test_gpu_sweep_parity)
            assert np.allclose(out3[i, 1:], one["out3"][1:], rtol=1e-9, atol=1e-12)
        else:
            assert np.allclose(out3[i], one["out3"], rtol=1e-8, atol=1e-12), (i, out3[i], one["out3"])


# gap ratios of the oracle (= the reference algorithm) for J = 1, U = 1..12, mu = 0.5; every one of these spectra was checked
# against dense diagonalisation on the CPU (tools/probe_multiplicity.py): the 20 lowest levels WITH their multiplicities
MULT_REF = {
    (6, 6): [0.12774755, 0.19702473, 0.04525686, 0.17866127, 0.19142464, 0.11412812, 0.15500726, 0.13156926, 0.14501948, 0.15547803,
             0.1615497, 0.17037436],
    (7, 5): [0.01978474, 0.04745109, 0.02946433, 0.00621864, 0.00318766, 0.01483454, 0.03836963, 0.05660714, 0.08258303, 0.07345474,
             0.05862871, 0.04995468],
    (5, 5): [0.0750343, 0.06344276, 0.08264736, 0.07231096, 0.06340804, 0.04345234, 0.01048562, 0.01403509, 0.01823648, 0.02117524,
             0.02330482, 0.02491944],
}


@pytest.mark.parametrize("m,n", sorted(MULT_REF))
def test_multiplicities_on_small_chains(pkg, ctx_factory, m, n):
    """A single-vector Krylov method finds the second copy of an exactly degenerate level (momentum +k / -k) only through
    rounding; a solver that converges before the copy has grown returns a wrong 20-level multiset.  Round 1's plain mode did
    that at m = n = 6, U = 5 (found by this table); since round 2 the Chebyshev-filtered solver -- which amplifies the copy
    by orders of magnitude per application -- is used down to D = 100.  Every solver path must reproduce the table."""
    ref = np.array(MULT_REF[(m, n)])
    U = np.arange(1.0, 13.0)
    ctx = ctx_factory(m, n)
    many, _ = ctx.points(np.ones(12), U, np.full(12, 0.5), kernel=pkg.capi.HV_MATRIX_FREE)
    assert np.allclose(many[:, 0], ref, rtol=0, atol=2e-8)
    for kernel in (pkg.capi.HV_MATRIX_FREE, pkg.capi.HV_STORED):
        one = np.array([ctx.point(1.0, u, 0.5, kernel=kernel)["out3"][0] for u in U])
        assert np.allclose(one, ref, rtol=0, atol=2e-8), (kernel, one - ref)
