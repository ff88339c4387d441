"""Probe (GPU box): H.v count / time of the thick-restart solver as a function of the Krylov size ncv."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
m = n = int(sys.argv[1]); U = float(sys.argv[2]); ncvs = [int(v) for v in sys.argv[3].split(",")]
ctx = pkg.Context(0).setup(m, n)
base = None
for ncv in ncvs:
    r = ctx.eigs(1.0, U, 1.0, nev=20, ncv=ncv, allow_noconv=True)
    if base is None: base = r["evals"]
    print(f"m={m} U={U} ncv={ncv}: nmatvec={r['nmatvec']} nrestart={r['nrestart']} nconv={r['nconv']} t={r['seconds']*1e3:.1f} ms maxdiff={np.abs(r['evals']-base).max():.1e}", flush=True)
