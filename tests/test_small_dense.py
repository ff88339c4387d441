"""CPU: the host-side symmetric eigensolver of the projected problems (csrc/small_dense.cpp, replaces Spectra's
TridiagEigen and the m x m EigenSolver call) against numpy.linalg.eigh -- random, arrowhead + tridiagonal (the shape after a
thick restart), degenerate, diagonal and zero matrices."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eig():
    out = os.path.join(tempfile.mkdtemp(prefix="bh_sd_"), "libsd.so")
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", out,
                           os.path.join(ROOT, "tests", "small_dense_wrap.cpp"),
                           os.path.join(ROOT, "bose-hubbard-phase-transition_b200", "csrc", "small_dense.cpp")])
    L = C.CDLL(out)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.wrap_sym_eig.argtypes = [C.c_int, dp, dp, dp]

    def run(a):
        n = a.shape[0]
        ev = np.empty(n)
        v = np.empty(n * n)
        L.wrap_sym_eig(n, np.ascontiguousarray(a.T).reshape(-1), ev, v)   # column-major in, column-major out
        return ev, v.reshape(n, n).T.copy()
    return run


def check(run, a):
    n = a.shape[0]
    ev, v = run(a)
    want = np.linalg.eigvalsh(a)
    scale = max(np.abs(a).max(), 1e-300) * max(n, 1)
    assert np.all(np.diff(ev) >= 0)
    assert np.abs(ev - want).max() <= 1e-13 * scale
    assert np.abs(a @ v - v * ev).max() <= 1e-13 * scale
    assert np.abs(v.T @ v - np.eye(n)).max() <= 1e-13 * max(n, 1)


def test_random_symmetric(eig):
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 5, 12, 20, 41, 64):
        for _ in range(5):
            a = rng.uniform(-1, 1, (n, n))
            check(eig, a + a.T)


def test_arrowhead_plus_tridiagonal(eig):
    rng = np.random.default_rng(1)
    for n, k in ((41, 25), (41, 20), (21, 8), (10, 9)):
        a = np.zeros((n, n))
        a[np.arange(k), np.arange(k)] = np.sort(rng.uniform(-50, 50, k))
        a[k, :k] = a[:k, k] = rng.uniform(-1e-3, 1e-3, k)       # coupling row of the thick restart
        for i in range(k, n):
            a[i, i] = rng.uniform(-50, 50)
            if i + 1 < n:
                a[i, i + 1] = a[i + 1, i] = rng.uniform(0.1, 5)
        check(eig, a)


def test_degenerate_diagonal_and_zero(eig):
    rng = np.random.default_rng(2)
    q, _ = np.linalg.qr(rng.normal(size=(12, 12)))
    d = np.array([1.0, 1.0, 1.0, 2.0, 2.0, 3.0, 3.0, 3.0, 3.0, -1.0, -1.0, 0.0])
    check(eig, (q * d) @ q.T)
    check(eig, np.diag(np.arange(7.0)))
    check(eig, np.zeros((6, 6)))
    check(eig, np.full((5, 5), 2.0))
