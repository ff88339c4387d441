"""CPU: the halo plan of the row-partitioned H.v (csrc/halo_plan.h, the host logic of csrc/dist.cu for N > 1 GPUs).

The flags a GPU rank would produce with k_mark_halo_chain are rebuilt here from the oracle's basis (every hop of every row of a
rank's slice whose source lies outside the slice marks its chunk); the planner must then give, for every rank count and chunk
size, receive lists that cover every remote source element, stay inside the owning peer's slice, and mirror the peers' send
lists range for range (that is what makes the grouped ncclSend / ncclRecv of the ranks match)."""
import os
import subprocess
import tempfile
from math import comb

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lex_rank(states, n):
    m = states.shape[1]
    R = n - np.cumsum(states, axis=1)
    r = np.zeros(len(states), dtype=np.int64)
    for q in range(m - 1):
        f = np.array([0] + [comb(x - 1 + m - 1 - q, m - 1 - q) for x in range(1, n + 2)], dtype=np.int64)
        r += f[R[:, q]]
    return r


def remote_sources(m, n, W, per):
    """per rank: sorted unique global indices of the sources of its rows' hops that lie outside its slice (closed chain)"""
    _, bas = O.basis(m, n, O.LEX)
    bas = bas.astype(np.int64)
    D = len(bas)
    assert (lex_rank(bas, n) == np.arange(D)).all()
    out = []
    for r in range(W):
        lo, hi = min(D, r * per), min(D, (r + 1) * per)
        S = bas[lo:hi]
        rem = []
        for q in range(m):
            a, b = q, (q + 1) % m
            for src, dst in ((a, b), (b, a)):
                ok = S[:, src] > 0
                T = S[ok].copy()
                T[:, src] -= 1
                T[:, dst] += 1
                t = lex_rank(T, n)
                rem.append(t[(t < lo) | (t >= hi)])
        out.append(np.unique(np.concatenate(rem)) if rem else np.zeros(0, dtype=np.int64))
    return D, out


@pytest.fixture(scope="module")
def planner():
    exe = os.path.join(tempfile.mkdtemp(prefix="bh_halo_"), "halo_plan_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "halo_plan_test.cpp")])
    return exe


@pytest.mark.parametrize("m,n,W,chunk,gap", [(6, 6, 2, 16, 8), (7, 7, 2, 64, 2), (8, 8, 4, 64, 8), (8, 8, 8, 32, 0), (9, 6, 3, 128, 4),
                                             (8, 8, 5, 4096, 8)])
def test_halo_plan_covers_every_remote_source_and_pairs_up(planner, m, n, W, chunk, gap):
    D = O.dimension(m, n)
    per = ((D + W - 1) // W + 31) // 32 * 32   # slice length exactly as csrc/ctx_basis.cu computes it
    D, remote = remote_sources(m, n, W, per)
    nchunks = (per * W + chunk - 1) // chunk
    flags = np.zeros((W, nchunks), dtype=np.uint8)
    for r in range(W):
        flags[r, np.unique(remote[r] // chunk)] = 1
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        flags.tofile(f.name)
        out = subprocess.run([planner, str(W), str(per), str(D), str(chunk), str(gap), f.name], capture_output=True, text=True, check=True).stdout
    recv = {r: [] for r in range(W)}
    send = {r: [] for r in range(W)}
    for line in out.splitlines():
        me, kind, peer, off, cnt = line.split()
        (recv if kind == "r" else send)[int(me)].append((int(peer), int(off), int(cnt)))
    for r in range(W):
        covered = np.zeros(D, dtype=bool)
        for peer, off, cnt in recv[r]:
            assert peer != r and cnt > 0
            assert per * peer <= off and off + cnt <= min(D, per * (peer + 1))      # inside the owner's slice
            covered[off:off + cnt] = True
        assert covered[remote[r]].all()                                             # every remote source arrives
        if gap == 0 and chunk <= 64:                                                 # no merging: nothing but flagged chunks travels
            assert covered.sum() <= len(np.unique(remote[r] // chunk)) * chunk
        # the receive list of (r <- p) is the send list of (p -> r), in the same order
        for p in range(W):
            if p != r:
                assert [(o, c) for (q, o, c) in recv[r] if q == p] == [(o, c) for (q, o, c) in send[p] if q == r]
    # what the all-gather form moves per rank vs what the plan moves
    moved = sum(c for r in range(W) for (_, _, c) in recv[r])
    assert moved <= (W - 1) * per * W
