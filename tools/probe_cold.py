"""GPU box: where does the time of a cold context go?  usage: probe_cold.py m n [npoints]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as g
pkg = g.load_package(); capi = pkg.capi
m, n = int(sys.argv[1]), int(sys.argv[2])
npts = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cJ = np.ones(npts); cU = np.array([3.0 + (i // 4) for i in range(npts)]); cmu = np.array([float(i % 4) for i in range(npts)])
for rep in range(2):
    for b in (1, 4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx = pkg.Context(0); ctx.setup(m, n); ctx.set_batch(b)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        ctx.points(cJ, cU, cmu, kernel=capi.HV_MATRIX_FREE)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        ctx.points(cJ, cU, cmu, kernel=capi.HV_MATRIX_FREE)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        ctx.close()
        torch.cuda.synchronize(); t4 = time.perf_counter()
        print(f"rep {rep} batch {b}: setup {t1 - t0:.3f} s, first {npts} points {t2 - t1:.3f} s, next {npts} points {t3 - t2:.3f} s, close {t4 - t3:.3f} s", flush=True)
