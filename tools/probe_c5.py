#!/usr/bin/env python3
"""C5 (m = n = 14, nev = 2, ncv = 12) solve time for the cut fractions given on the command line (BH_CHEB_FRAC, read at
context creation): three timed solves each after a warm-up."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
capi = pkg.capi
for frac in (sys.argv[1:] or ["0.05", "0.08"]):
    os.environ["BH_CHEB_FRAC"] = frac
    ctx = pkg.Context(0).setup(14, 14)
    pars = (1.0, 4.0, 1.0)
    ctx.eigs(*pars, nev=2, ncv=12, kernel=capi.HV_MATRIX_FREE, order=capi.LEX, maxit=2, allow_noconv=True)
    for _ in range(3):
        r = ctx.eigs(*pars, nev=2, ncv=12, kernel=capi.HV_MATRIX_FREE, order=capi.LEX)
        print(json.dumps({"frac": frac, "solve_s": r["seconds"], "nmatvec": r["nmatvec"], "nrestart": r.get("nrestart"),
                          "nreorth": r.get("nreorth"), "E0": float(r["evals"][0])}), flush=True)
    ctx.close()
