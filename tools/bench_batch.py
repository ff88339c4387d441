"""GPU box: points/s of bh_points with lockstep batching (batch = 1, 2, 4).  usage: bench_batch.py m n npoints [batches]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as g
pkg = g.load_package(); capi = pkg.capi
m, n, npts = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batches = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 4]
uperm = np.random.default_rng(0).permutation(32)
cU = np.array([1.0 + uperm[i % 32] for i in range(npts)])
cmu = np.array([float(i % 32) for i in range(npts)])
cJ = np.ones(npts)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = pkg.Context(0); ctx.set_stream(stream.cuda_stream); ctx.setup(m, n)
ref = None
for b in batches:
    ctx.set_batch(b)
    ctx.points(cJ[:max(b, 2)], cU[:max(b, 2)], cmu[:max(b, 2)], kernel=capi.HV_MATRIX_FREE)  # warm-up: workspaces
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    out3, infos = ctx.points(cJ, cU, cmu, kernel=capi.HV_MATRIX_FREE)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if ref is None: ref = out3
    print(f"m={m} n={n} batch={b}: {npts / dt:.2f} points/s ({dt / npts * 1e3:.1f} ms/point), launches {ctx.launch_count() - l0}, "
          f"H.v/point {np.mean([i['nmatvec'] for i in infos]):.0f}, max |d out3| vs batch=1 {np.abs(out3 - ref).max():.1e}", flush=True)
ctx.close()
