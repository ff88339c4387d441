"""Probe (GPU box): time per grid point of the accelerated solver for a list of U values (warm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
m = n = int(sys.argv[1]); Us = [float(u) for u in sys.argv[2].split(",")]
ctx = pkg.Context(0).setup(m, n)
ctx.eigs(1.0, 2.0, 0.0, nev=20, maxit=2, allow_noconv=True)
tot = 0; tmv = 0
for U in Us:
    r = ctx.eigs(1.0, U, 1.0, nev=20, kernel=1, allow_noconv=True)
    tot += r["seconds"]; tmv += r["nmatvec"]
    print(f"  U={U}: {r['seconds']*1e3:.1f} ms nmatvec={r['nmatvec']} nrestart={r['nrestart']} nconv={r['nconv']} E19={r['evals'][19]:.9f}", flush=True)
print(f"m={m} env={ {k:v for k,v in os.environ.items() if k.startswith('BH_')} } total {tot*1e3:.1f} ms, {len(Us)/tot:.2f} points/s, mean nmatvec {tmv/len(Us):.0f}", flush=True)
