#!/usr/bin/env python3
"""Small drivers for `ncu --set full` captures (one kernel each; see tools/gpu_round2_k.sh):
   c3  one lockstep list of 4 grid points at m = n = 12 (k_step_coop, k_hv_chain_batch, k_compress_tiled8)
   c2  32 grid points at m = n = 10 through the many-point small-system solver (k_small_cycle)
   c5  matrix-free H.v at m = n = 14 (k_hv_free_chain<14>)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
mode = sys.argv[1]
if mode == "c3":
    ctx = pkg.Context(0).setup(12, 12)
    ctx.set_batch(4)
    U = np.array([3.0, 9.0, 17.0, 26.0])
    ctx.points(np.ones(4), U, np.zeros(4), kernel=pkg.capi.HV_MATRIX_FREE)
elif mode == "c2":
    ctx = pkg.Context(0).setup(10, 10)
    U = 1.0 + np.arange(32.0) % 11
    ctx.points(np.ones(32), U, np.arange(32.0) % 7, kernel=pkg.capi.HV_MATRIX_FREE)
elif mode == "c5":
    import torch
    ctx = pkg.Context(0).setup(14, 14)
    x = torch.zeros(ctx.D, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
    ctx.lcg_fill_dev(x.data_ptr(), ctx.D)
    for _ in range(3):
        ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), pkg.capi.HV_MATRIX_FREE)
    torch.cuda.synchronize()
ctx.close()
