"""Probe (GPU box): H.v counts / times of the GPU solver vs the CPU restatement and the compiled reference."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import oracle_lib as O, ref_lib as R
pkg = g.load_package()
m = n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Us = [float(u) for u in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 10, 32]
use_ref = len(sys.argv) > 3 and sys.argv[3] == "ref"
ctx = pkg.Context(0).setup(m, n)
if not use_ref:
    t, b = O.basis(m, n); jc = O.hopping_csc(m, O.chain(m), t, b); dU, dN = O.diagonals(m, b)
for U in Us:
    for kern in (0, 1):
        r = ctx.eigs(1.0, U, 1.0, kernel=kern, allow_noconv=True)
        print(f"m={m} U={U} kernel={kern}: gpu nmatvec={r['nmatvec']} nrestart={r['nrestart']} nreorth={r['nreorth']} nconv={r['nconv']} t={r['seconds']*1e3:.1f} ms E0={r['evals'][0]:.12f}", flush=True)
    if use_ref:
        rr, info = R.eigs(m, n, 1, U, 1)
        print(f"     ref nmatvec={info['nmatvec']} nrestart={info['nrestart']} t={info['seconds']:.2f}s maxdiff={np.abs(np.sort(rr['evals'])-r['evals']).max():.2e}", flush=True)
    else:
        h = O.hsum_csc(jc, dU, dN, 1.0, U, 1.0)
        t0 = time.time(); o = O.eigs_sym(h); t1 = time.time()
        print(f"     oracle nmatvec={o['nmatvec']} nrestart={o['nrestart']} t={t1-t0:.2f}s maxdiff={np.abs(o['evals']-r['evals']).max():.2e}", flush=True)
