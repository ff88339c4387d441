#!/usr/bin/env python3
"""CPU model (numpy/scipy) of the eigensolver variants, used to choose the algorithm before writing kernels:
thick-restart *block* Lanczos with full re-orthogonalisation on a Chebyshev filter of H, block size b
(b = 1 is the single-vector scheme of csrc/lanczos.cu).  Counts H.v applications until Spectra's stopping
rule holds for the nev lowest Ritz pairs of the filtered operator, then Rayleigh-Ritz of H.

    python tools/model_block_lanczos.py M U[,U..] b[,b..] ncv[,ncv..] d[,d..]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402


def build_H(m, n):
    tags, bas = O.basis(m, n, O.TAG_SORTED)
    o, i, v = O.hopping_csc(m, O.chain(m), tags, bas)
    D = len(tags)
    JH = sp.csc_matrix((v, i, o), shape=(D, D)).tocsr()
    dU, dN = O.diagonals(m, bas)
    return JH, dU, dN


class Counter:
    def __init__(self, H):
        self.H = H
        self.n = 0

    def __call__(self, X):
        self.n += 1 if X.ndim == 1 else X.shape[1]
        return self.H @ X


def cheb_op(hv, c, e, d):
    def op(X):
        t0 = X
        t1 = (hv(X) - c * X) / e
        for _ in range(2, d + 1):
            t2 = 2.0 * (hv(t1) - c * t1) / e - t0
            t0, t1 = t1, t2
        return -t1 if d % 2 == 0 else t1
    return op


def tr_block_lanczos(op, X0, nev, ncv, tol, maxit, b, keep_rule="spectra"):
    """-> (theta[nev], V, Y, nrestart, converged)"""
    D = X0.shape[0]
    V = np.zeros((D, ncv + b))
    T = np.zeros((ncv + b, ncv + b))
    Q, _ = np.linalg.qr(X0)
    V[:, :b] = Q
    cols = b
    eps23 = np.finfo(float).eps ** (2.0 / 3)
    it = 0
    while True:
        while cols <= ncv:
            W = op(V[:, cols - b:cols])
            C = V[:, :cols].T @ W
            W -= V[:, :cols] @ C
            C2 = V[:, :cols].T @ W
            W -= V[:, :cols] @ C2
            T[:cols, cols - b:cols] = C + C2
            Qn, R = np.linalg.qr(W)
            V[:, cols:cols + b] = Qn
            T[cols:cols + b, cols - b:cols] = R
            cols += b
        n = cols - b  # basis size used for Rayleigh-Ritz (== ncv when (ncv - k) % b == 0)
        Tn = 0.5 * (T[:n, :n] + T[:n, :n].T)
        th, Y = np.linalg.eigh(Tn)
        Rl = T[n:n + b, n - b:n].copy()  # coupling of the next block
        res = np.linalg.norm(Rl @ Y[n - b:n, :], axis=0)
        nconv = int(np.sum(res[:nev] < tol * np.maximum(eps23, np.abs(th[:nev]))))
        if nconv >= nev or it >= maxit:
            return th[:nev], V[:, :n], Y, it + 1, nconv >= nev
        it += 1
        k = nev + min(nconv, (n - nev) // 2)
        while (ncv - k) % b:
            k += 1
        k = min(k, n - b)
        Vk = V[:, :n] @ Y[:, :k]
        Qn = V[:, n:n + b].copy()
        V[:, :k] = Vk
        V[:, k:k + b] = Qn
        T[:] = 0.0
        T[:k, :k] = np.diag(th[:k])
        S = Rl @ Y[n - b:n, :k]
        T[k:k + b, :k] = S
        T[:k, k:k + b] = S.T
        cols = k + b


def solve(JH, dU, n, cU, cmu, b, ncv, d, nev=20, tol=1e-10, pre=3, margin=0.05, frac=0.08, pre_ncv=41):
    D = JH.shape[0]
    H = (JH + sp.diags(cU * dU - cmu * n)).tocsr()
    hv = Counter(H)
    rng = np.random.default_rng(0)
    # stage 1: plain cycles of the single-vector scheme (as csrc/lanczos.cu)
    x0 = hv(rng.uniform(-0.5, 0.5, (D, 1)))
    th1, V1, Y1, _, ok = tr_block_lanczos(hv, x0, nev, pre_ncv, tol, pre, 1)
    n1 = hv.n
    if ok:
        return hv.n, n1, 0, th1
    absrow = np.abs(H).sum(axis=1).A1 if hasattr(np.abs(H).sum(axis=1), "A1") else np.asarray(np.abs(H).sum(axis=1)).ravel()
    hi = float(np.max(2 * H.diagonal() - 0 + (absrow - np.abs(H.diagonal())) - H.diagonal()))  # diag + offdiag sum
    hi = float(np.max(H.diagonal() + (absrow - np.abs(H.diagonal()))))
    hi += 1e-9 * abs(hi) + 1e-12
    cut = max(th1[nev - 1] + margin * (th1[nev - 1] - th1[0]), th1[0] + frac * (hi - th1[0]))
    c, e = 0.5 * (hi + cut), 0.5 * (hi - cut)
    op = cheb_op(hv, c, e, d)
    # start block: b combinations of the stage-1 Ritz vectors
    Rz = V1 @ Y1[:, :nev]
    if b == 1:
        X0 = Rz.sum(axis=1, keepdims=True)
    else:
        X0 = Rz @ rng.standard_normal((nev, b))
    th2, V2, Y2, nr, ok = tr_block_lanczos(op, X0, nev, ncv, tol, 1000, b)
    n2 = hv.n - n1
    # stage 3: Rayleigh-Ritz of H
    W = hv(V2)
    M = V2.T @ W
    ev = np.linalg.eigvalsh(0.5 * (M + M.T))
    return hv.n, n1, nr, ev[:nev]


if __name__ == "__main__":
    m = int(sys.argv[1])
    Us = [float(u) for u in sys.argv[2].split(",")]
    bs = [int(x) for x in sys.argv[3].split(",")]
    ncvs = [int(x) for x in sys.argv[4].split(",")]
    ds = [int(x) for x in sys.argv[5].split(",")]
    JH, dU, dN = build_H(m, m)
    import scipy.sparse.linalg as sla
    for U in Us:
        H = (JH + sp.diags(U * dU - 1.0 * m)).tocsr()
        ref = None
        if JH.shape[0] < 20000:
            ref = np.sort(sla.eigsh(H, k=20, which="SA", tol=1e-13)[0])
        for b in bs:
            for ncv in ncvs:
                for d in ds:
                    tot, n1, nr, ev = solve(JH, dU, m, U, 1.0, b, ncv, d)
                    err = np.max(np.abs(ev - ref)) if ref is not None else float("nan")
                    print(f"m={m} U={U:g} b={b} ncv={ncv} d={d}: H.v total {tot} (stage1 {n1}) filter-restarts {nr} "
                          f"filter-applications/vector {(tot - n1 - ncv) // d}  err {err:.2e}", flush=True)
