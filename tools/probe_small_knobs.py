#!/usr/bin/env python3
"""Sweep of the Chebyshev-filter knobs (BH_CHEB_FRAC here; BH_CHEB_DEGREE and BH_CHEB_QUICK were swept the same way in session Y; read at context creation) on the C2 sweep
(m = n = 10, 121 points, many-point solver) and on 16 points of C3 (m = n = 12, lockstep batches of 4).  Prints points/s and the
largest deviation of the three output columns from the default setting's."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()


def run(m, cJ, cU, cmu, env, reps=2):
    for k in ("BH_CHEB_DEGREE", "BH_CHEB_FRAC", "BH_CHEB_QUICK", "BH_CHEB_MARGIN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = pkg.Context(0).setup(m, m)
    ctx.set_batch(4)
    ctx.points(cJ, cU, cmu, kernel=pkg.capi.HV_MATRIX_FREE)
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        out3, infos = ctx.points(cJ, cU, cmu, kernel=pkg.capi.HV_MATRIX_FREE)
        best = min(best, time.perf_counter() - t)
    ctx.close()
    return len(cU) / best, out3, sum(i["nmatvec"] for i in infos) / len(cU)


which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c2":
    U, MU = np.meshgrid(1.0 + np.arange(11), np.arange(11.0), indexing="ij")
    cU, cmu = U.ravel(), MU.ravel()
    m = 10
else:
    cU = 2.0 * (1 + np.arange(16)); cmu = np.zeros(16)
    m = 12
cJ = np.ones(len(cU))
base_v, base, base_mv = run(m, cJ, cU, cmu, {})
print(json.dumps({"cfg": which, "env": {}, "points_per_s": base_v, "mean_matvec": base_mv}), flush=True)
fracs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0.04", "0.06", "0.08"]
grid = [{"BH_CHEB_FRAC": f} for f in fracs]
for env in grid:
    try:
        v, out3, mv = run(m, cJ, cU, cmu, env)
        print(json.dumps({"cfg": which, "env": env, "points_per_s": v, "mean_matvec": mv,
                          "max_dev": float(np.abs(out3 - base).max())}), flush=True)
    except Exception as ex:
        print(json.dumps({"cfg": which, "env": env, "error": str(ex)[:200]}), flush=True)
