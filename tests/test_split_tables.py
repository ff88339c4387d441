"""CPU check of the split matrix-free H.v kernel's tables and work decomposition (hv_split_tables.h).

tests/split_emul.cpp replays the kernel's loops (items -> warps -> lanes -> prefixes) on the host tables; the result
must equal the oracle's H.x and every row must be written exactly once, for every cut position p and group size G."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(tempfile.mkdtemp(prefix="bh_split_emul_"), "libsplit_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "split_emul.cpp")])
    L = C.CDLL(out)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.split_emul_hv.argtypes = [C.c_int] * 5 + [C.c_double] * 3 + [C.c_long, dp, dp, ip]
    return L


def lex_problem(m, n, closed):
    """oracle H pieces in the tag-sorted order + the permutation LEX position -> tag position"""
    tags, bas = O.basis(m, n, O.TAG_SORTED)
    _, lbas = O.basis(m, n, O.LEX)
    pos = {tuple(r): i for i, r in enumerate(bas.astype(int).tolist())}
    perm = np.array([pos[tuple(r)] for r in lbas.astype(int).tolist()])
    jc = O.hopping_csc(m, O.chain(m, closed), tags, bas, 1.0)
    dU, dN = O.diagonals(m, bas)
    return jc, dU, dN, perm


@pytest.mark.parametrize("m,n", [(3, 3), (3, 5), (4, 4), (5, 5), (6, 6), (7, 5), (5, 9), (8, 8)])
@pytest.mark.parametrize("closed", [True, False])
def test_split_tables_reproduce_oracle_hv(emul, m, n, closed):
    jc, dU, dN, perm = lex_problem(m, n, closed)
    D = len(perm)
    xt = O.lcg_vector(D)
    x = np.ascontiguousarray(xt[perm])
    for (cJ, cU, cmu) in [(1.0, 4.0, 1.0), (0.3, 0.0, 2.0)]:
        want = O.spmv(O.hsum_csc(jc, dU, dN, cJ, cU, cmu), xt)[perm]
        for p in range(1, m):
            for G in ((1, 8) if D > 1000 else (1, 4, 8, 16)):
                y = np.zeros(D)
                touched = np.zeros(D, dtype=np.int32)
                rc = emul.split_emul_hv(m, n, p, G, int(closed), cJ, cU, cmu, D, x, y, touched)
                assert rc == 0, (rc, p, G)
                assert (touched == 1).all(), (p, G, touched.min(), touched.max())
                assert np.abs(y - want).max() <= 1e-13 * np.abs(want).max(), (p, G)
