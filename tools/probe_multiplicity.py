"""Probe (GPU box): do the solver paths return the true 20 lowest levels WITH multiplicities on small closed chains?
Reference values: gap ratios of the oracle (= the reference algorithm), verified against dense diagonalisation on the CPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
REF = {
 (6, 6): [0.12774755, 0.19702473, 0.04525686, 0.17866127, 0.19142464, 0.11412812, 0.15500726, 0.13156926, 0.14501948, 0.15547803, 0.1615497, 0.17037436],
 (7, 5): [0.01978474, 0.04745109, 0.02946433, 0.00621864, 0.00318766, 0.01483454, 0.03836963, 0.05660714, 0.08258303, 0.07345474, 0.05862871, 0.04995468],
 (5, 5): [0.0750343, 0.06344276, 0.08264736, 0.07231096, 0.06340804, 0.04345234, 0.01048562, 0.01403509, 0.01823648, 0.02117524, 0.02330482, 0.02491944],
}
U = np.arange(1.0, 13.0)
for (m, n), ref in REF.items():
    ctx = pkg.Context(0).setup(m, n)
    ref = np.array(ref)
    small, _ = ctx.points(np.ones(12), U, np.full(12, 0.5), kernel=pkg.capi.HV_MATRIX_FREE)
    single_f = np.array([ctx.point(1.0, u, 0.5, kernel=pkg.capi.HV_MATRIX_FREE)["out3"][0] for u in U])
    single_s = np.array([ctx.point(1.0, u, 0.5, kernel=pkg.capi.HV_STORED)["out3"][0] for u in U])
    bad = lambda a: [int(u) for u, x, r in zip(U, a, ref) if abs(x - r) > 1e-6]
    print(f"m={m} n={n} env={ {k: v for k, v in os.environ.items() if k.startswith('BH_')} }: wrong U: small path {bad(small[:, 0])}, "
          f"single matrix-free {bad(single_f)}, single stored {bad(single_s)}", flush=True)
    ctx.close()
