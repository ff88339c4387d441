"""GPU box: where does the time of a cold context go?  usage: probe_cold.py m n"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as g
pkg = g.load_package(); capi = pkg.capi
m, n = int(sys.argv[1]), int(sys.argv[2])
cJ = np.ones(2); cU = np.array([3.0, 17.0]); cmu = np.array([1.0, 2.0])
for rep in range(2):
    for b in (1, 2):
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
        print(f"rep {rep} batch {b}: setup {t1 - t0:.3f} s, first 2 points {t2 - t1:.3f} s, next 2 points {t3 - t2:.3f} s, close {t4 - t3:.3f} s", flush=True)
