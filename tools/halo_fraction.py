#!/usr/bin/env python3
"""Hop-graph analysis of the row partition (CPU, numpy): for W contiguous LEX slices of a closed chain, the fraction of D that a
rank reads from the other slices, the remote hops per row, and how the halo splits over P consecutive row pieces of a slice (the
numbers quoted in DESIGN.md section 7).

    python tools/halo_fraction.py M W [P]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from test_halo_plan import lex_rank  # noqa: E402

m = n = int(sys.argv[1])
W = int(sys.argv[2])
P = int(sys.argv[3]) if len(sys.argv) > 3 else 4
_, bas = O.basis(m, n, O.LEX)
bas = bas.astype(np.int64)
D = len(bas)
per = -(-D // W)
CH = max(1, D // 2000)
for r in range(W):
    lo, hi = r * per, min(D, (r + 1) * per)
    S = bas[lo:hi]
    piece = (np.arange(hi - lo) * P) // (hi - lo)
    need = [set() for _ in range(P)]
    rem_hops = 0
    for q in range(m):
        a, b = q, (q + 1) % m
        for src, dst in ((a, b), (b, a)):
            ok = S[:, src] > 0
            T = S[ok].copy()
            T[:, src] -= 1
            T[:, dst] += 1
            t = lex_rank(T, n)
            remote = (t < lo) | (t >= hi)
            rem_hops += remote.sum()
            pc, ch = piece[ok][remote], t[remote] // CH
            for p in range(P):
                need[p].update(np.unique(ch[pc == p]).tolist())
    seen, inc = set(), []
    for p in range(P):
        inc.append(len(need[p] - seen))
        seen |= need[p]
    print(f"rank {r}/{W}: rows {hi - lo}, remote hops per row {rem_hops / (hi - lo):.2f}, halo {len(seen) * CH / D:.3f} D "
          f"(all-gather: {(W - 1) / W:.3f} D); new chunks per row piece {inc}")
