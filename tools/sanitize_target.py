#!/usr/bin/env python3
"""Workload for compute-sanitizer: the smoke point (m = n = 6, both H.v kernels: cooperative Lanczos
step, SELL build, SPDM) and a lockstep-batched bh_points call at m = n = 8 (baton scheduler, interleaved batch kernels,
double-buffered partial sums of the cooperative step).

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
ctx = pkg.Context(0).setup(6, 6)
for kernel in (pkg.capi.HV_STORED, pkg.capi.HV_MATRIX_FREE):
    r = ctx.point(1.0, 4.0, 1.0, kernel=kernel)
    print("m=6 point kernel", kernel, r["out3"], flush=True)
ctx.close()
npts = int(os.environ.get("SAN_POINTS", "6"))
ctx = pkg.Context(0).setup(8, 8)
ctx.set_batch(4)
U = 1.0 + np.arange(npts)
out3, infos = ctx.points(np.ones(npts), U, np.zeros(npts), kernel=pkg.capi.HV_MATRIX_FREE)
print("m=8 lockstep points", out3[:, 0], [i["nmatvec"] for i in infos], flush=True)
ctx.close()
print("sanitize target done", flush=True)
