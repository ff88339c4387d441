"""Probe (GPU box): warm per-step cost of the Lanczos solver.  usage: probe_step.py m n U ncv [maxit]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
m, n, U, ncv = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
maxit = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
ctx = pkg.Context(0).setup(m, n)
ctx.eigs(1.0, U, 1.0, nev=20, ncv=ncv, maxit=2, allow_noconv=True)   # warm-up: allocations, SELL build
for kern in (0, 1):
    r = ctx.eigs(1.0, U, 1.0, nev=20, ncv=ncv, maxit=maxit, kernel=kern, allow_noconv=True)
    print(f"m={m} U={U} ncv={ncv} kernel={kern}: nmatvec={r['nmatvec']} nrestart={r['nrestart']} nconv={r['nconv']} "
          f"t={r['seconds']*1e3:.1f} ms -> {r['seconds']*1e6/r['nmatvec']:.1f} us/step", flush=True)
