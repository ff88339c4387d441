#!/usr/bin/env python3
"""CPU model of where to place the filter's cut (csrc/lanczos.cu accel_solve, quick stage 1): H.v count of
  fixed     cut = theta_0 + frac * (hi - theta_0)                                   (the round-2 rule, frac = 0.08)
  explore   the same cut for ONE restart cycle of the filtered iteration, whose Ritz values (upper bounds of E_0..E_19 after
            inverting the Chebyshev polynomial) place a tighter cut  E19' + marg * (E19' - E0'); restart from the Ritz vectors' sum
  ideal     cut = E_0 + (1 + marg) * (E_19 - E_0) with the true levels (what a perfect estimate would give)

    python tools/model_cut.py M U[,U..]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from model_block_lanczos import Counter, build_H, cheb_op, tr_block_lanczos


def gersh_hi(H):
    a = abs(H)
    off = np.asarray(a.sum(axis=1)).ravel() - np.abs(H.diagonal())
    hi = float(np.max(H.diagonal() + off))
    lo = float(np.min(H.diagonal() - off))
    return hi + 1e-9 * (hi - lo) + 1e-12


def invert(mu, c, e, d):
    """E with |T_d((E - c) / e)| = mu on the branch below the damped interval."""
    mu = np.maximum(np.abs(mu), 1.0)
    return c - e * np.cosh(np.arccosh(mu) / d)


def stage2(hv, x0, cut, hi, d, nev, ncv, maxit):
    c, e = 0.5 * (hi + cut), 0.5 * (hi - cut)
    op = cheb_op(hv, c, e, d)
    th, V, Y, nr, ok = tr_block_lanczos(op, x0, nev, ncv, 1e-10, maxit, 1)
    return th, V, Y, nr, ok, c, e


def solve(H, mode, d=8, nev=20, ncv=41, frac=0.08, marg=0.3, quick=24, explore_ncv=41, truth=None):
    D = H.shape[0]
    hv = Counter(H)
    rng = np.random.default_rng(0)
    hi = gersh_hi(H)
    x0 = hv(rng.uniform(-0.5, 0.5, (D, 1)))
    th1, V1, Y1, _, _ = tr_block_lanczos(hv, x0, 1, quick, 1e-10, 0, 1)
    th0 = th1[0]
    start = V1 @ Y1[:, :1]
    cut = th0 + frac * (hi - th0)
    n_explore = 0
    if mode == "ideal":
        cut = truth[0] + (1 + marg) * (truth[nev - 1] - truth[0])
    elif mode == "explore":
        before = hv.n
        th, V, Y, nr, ok, c, e = stage2(hv, start, cut, hi, d, nev, explore_ncv, 0)
        n_explore = hv.n - before
        # th ascending of -T_d (even d): the wanted end is the most negative
        est = np.sort(invert(th[:nev], c, e, d))
        cut2 = est[nev - 1] + marg * (est[nev - 1] - est[0])
        if cut2 < cut:
            cut = cut2
            start = (V @ Y[:, :nev]).sum(axis=1, keepdims=True)
    th, V, Y, nr, ok, c, e = stage2(hv, start, cut, hi, d, nev, ncv, 80)
    W = hv(V)
    M = V.T @ W
    ev = np.linalg.eigvalsh(0.5 * (M + M.T))[:nev]
    return hv.n, n_explore, nr, ok, ev, cut


if __name__ == "__main__":
    m = int(sys.argv[1])
    Us = [float(u) for u in sys.argv[2].split(",")]
    JH, dU, dN = build_H(m, m)
    import scipy.sparse.linalg as sla
    for U in Us:
        H = (JH + sp.diags(U * dU)).tocsr()
        truth = np.sort(sla.eigsh(H, k=24, which="SA", tol=1e-12, ncv=80)[0])
        S = truth[19] - truth[0]
        W = gersh_hi(H) - truth[0]
        print(f"m={m} U={U:g}: E0 {truth[0]:.6f} spread S {S:.4f} width W {W:.2f} S/W {S / W:.4f}", flush=True)
        for mode, kw in [("fixed", {"frac": 0.08}), ("fixed", {"frac": 0.06}), ("fixed", {"frac": 0.04}),
                         ("explore", {"marg": 0.3}), ("explore", {"marg": 0.6}), ("explore", {"marg": 0.3, "explore_ncv": 30}),
                         ("ideal", {"marg": 0.3}), ("ideal", {"marg": 0.6}), ("ideal", {"marg": 1.0})]:
            tot, nex, nr, ok, ev, cut = solve(H, mode, truth=truth, **kw)
            err = np.max(np.abs(ev - truth[:20]))
            print(f"   {mode:8s} {kw}: H.v {tot} (explore {nex}) restarts {nr} ok {ok} cut-E0 {(cut - truth[0]) / S:.2f} S  err {err:.1e}", flush=True)
