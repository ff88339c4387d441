"""Driver for ncu launch lists (GPU box): one warm grid point.  usage: probe_point.py m n U [kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
m, n, U = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
kern = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ctx = pkg.Context(0).setup(m, n)
ctx.eigs(1.0, 2.0, 0.0, nev=20, maxit=1, kernel=kern, allow_noconv=True)
r = ctx.point(1.0, U, 1.0, kernel=kern)
print(f"point m={m} U={U}: {r['seconds']*1e3:.1f} ms nmatvec={r['nmatvec']} nrestart={r['nrestart']} out3={r['out3']}")
