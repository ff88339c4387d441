#!/usr/bin/env python3
"""The quick-mode cut placed below the 20th level (small U: the 20 levels span more than 0.05 of the spectrum): the early stall exit
must hand the point to the full form with the right answer.  m = n = 10, a list through the many-point solver (csrc/small.cu) and the same
points one at a time (csrc/lanczos.cu), against the plain reference algorithm (BH_CHEB_DEGREE=1) on the same GPU."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
capi = pkg.capi
U = np.array([0.05, 0.1, 0.2, 0.3, 1.0, 2.0, 4.0, 8.0])
os.environ["BH_BATCH_VERBOSE"] = "1"
ctx = pkg.Context(0).setup(10, 10)
out3, infos = ctx.points(np.ones(len(U)), U, np.zeros(len(U)), kernel=capi.HV_MATRIX_FREE)
acc = [ctx.eigs(1.0, u, 0.0, nev=20, kernel=capi.HV_MATRIX_FREE, order=capi.LEX) for u in U[:4]]
ctx.close()
os.environ["BH_CHEB_DEGREE"] = "1"
ctx = pkg.Context(0).setup(10, 10)
for i, u in enumerate(U[:4]):
    ref = ctx.eigs(1.0, u, 0.0, nev=20, kernel=capi.HV_MATRIX_FREE, order=capi.LEX)
    one = ctx.point(1.0, u, 0.0, kernel=capi.HV_MATRIX_FREE)
    de = float(np.abs(np.sort(acc[i]["evals"]) - np.sort(ref["evals"])).max())
    print(json.dumps({"U": float(u), "list_nmatvec": infos[i]["nmatvec"], "single_nmatvec": acc[i]["nmatvec"], "single_s": acc[i]["seconds"],
                      "plain_nmatvec": ref["nmatvec"], "max_dE_vs_plain": de,
                      "out3_list": out3[i].tolist(), "out3_plain": np.asarray(one["out3"]).tolist()}), flush=True)
ctx.close()
