"""ctypes/numpy wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference (oracle/bh_oracle.cpp).  It is
the checker for the CUDA path; nothing under bose-hubbard-phase-transition_b200/
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

LEX, TAG_SORTED, REF_SCATTER = 0, 1, 2

_lib = None


def build():
    src = os.path.join(ORACLE_DIR, "bh_oracle.cpp")
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        L.bho_binomial.restype = C.c_int
        L.bho_dimension.restype = C.c_int
        L.bho_basis.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp]
        L.bho_search_tag.argtypes = [dp, C.c_int, C.c_double, C.c_double]
        L.bho_hopping_csc.restype = C.c_long
        L.bho_hopping_csc.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bho_diagonals.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.bho_hsum_csc.restype = C.c_long
        L.bho_hsum_csc.argtypes = [C.c_int, ip, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
        L.bho_spmv_csc.argtypes = [C.c_int, ip, ip, dp, dp, dp]
        L.bho_lcg_vector.argtypes = [C.c_long, dp]
        L.bho_eigs_sym.argtypes = [C.c_int, ip, ip, dp, C.c_int, C.c_int, C.c_double, C.c_int, dp, C.c_void_p,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.bho_dense_sym_eig.argtypes = [C.c_int, dp, dp, C.c_void_p]
        L.bho_gap_ratios.argtypes = [dp, C.c_int, dp]
        L.bho_spdm.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_int, dp]
        L.bho_condensate_fraction.restype = C.c_double
        L.bho_condensate_fraction.argtypes = [C.c_int, dp]
        L.bho_coherence.restype = C.c_double
        L.bho_coherence.argtypes = [C.c_int, dp]
        L.bho_point.argtypes = [C.c_int, C.c_int, dp, dp, ip, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                C.c_int, dp, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def dimension(m, n):
    return lib().bho_dimension(m, n)


def basis(m, n, order=TAG_SORTED):
    """-> tags[D], basis[D, m] (row k = state k; the reference's MatrixXd(m, D) column-major)."""
    D = dimension(m, n)
    tags = np.empty(D)
    bas = np.empty((D, m))
    rc = lib().bho_basis(m, n, order, tags, bas)
    assert rc == 0
    return tags, bas


def search_tag(tags, x, tol=1e-12):
    return lib().bho_search_tag(tags, len(tags), x, tol)


def chain(m, closed=True):
    """Neighbour list of src/neighbours.cpp:21-34 as (ptr, idx)."""
    nei = [[] for _ in range(m)]
    for i in range(m):
        if i > 0:
            nei[i].append(i - 1)
        if i < m - 1:
            nei[i].append(i + 1)
    if closed:
        nei[0].append(m - 1)
        nei[m - 1].append(0)
    return to_csr(nei)


def rect(lx, ly):
    """Periodic lx x ly rectangle, site = y*lx + x, order left,right,up,down (config 4)."""
    nei = []
    for y in range(ly):
        for x in range(lx):
            nei.append([y * lx + (x - 1) % lx, y * lx + (x + 1) % lx, ((y - 1) % ly) * lx + x, ((y + 1) % ly) * lx + x])
    return to_csr(nei)


def to_csr(nei):
    ptr = np.zeros(len(nei) + 1, dtype=np.int32)
    for i, l in enumerate(nei):
        ptr[i + 1] = ptr[i] + len(l)
    idx = np.array([s for l in nei for s in l], dtype=np.int32)
    return ptr, idx


def hopping_csc(m, nbr, tags, bas, J=1.0):
    D = len(tags)
    ptr, idx = nbr
    nnz = lib().bho_hopping_csc(m, D, ptr, idx, tags, bas, J, None, None, None)
    outer = np.empty(D + 1, dtype=np.int32)
    inner = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz)
    lib().bho_hopping_csc(m, D, ptr, idx, tags, bas, J, outer.ctypes.data, inner.ctypes.data, val.ctypes.data)
    return outer, inner, val


def diagonals(m, bas):
    D = bas.shape[0]
    dU = np.empty(D)
    dN = np.empty(D)
    lib().bho_diagonals(m, D, bas, dU, dN)
    return dU, dN


def hsum_csc(jcsc, dU, dN, cJ, cU, cu):
    outer_j, inner_j, val_j = jcsc
    D = len(dU)
    nnz = lib().bho_hsum_csc(D, outer_j, inner_j, val_j, dU, dN, cJ, cU, cu, None, None, None)
    outer = np.empty(D + 1, dtype=np.int32)
    inner = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz)
    lib().bho_hsum_csc(D, outer_j, inner_j, val_j, dU, dN, cJ, cU, cu, outer.ctypes.data, inner.ctypes.data,
                       val.ctypes.data)
    return outer, inner, val


def spmv(csc, x):
    outer, inner, val = csc
    y = np.empty(len(x))
    lib().bho_spmv_csc(len(x), outer, inner, val, np.ascontiguousarray(x), y)
    return y


def lcg_vector(n):
    out = np.empty(n)
    lib().bho_lcg_vector(n, out)
    return out


def eigs_sym(csc, nev=20, ncv=None, tol=1e-10, maxit=1000, want_vectors=False):
    outer, inner, val = csc
    D = len(outer) - 1
    ncv = ncv or 2 * nev + 1
    evals = np.empty(nev)
    vecs = np.empty((nev, D)) if want_vectors else None
    nmv, nrs = C.c_int(0), C.c_int(0)
    nconv = lib().bho_eigs_sym(D, outer, inner, val, nev, ncv, tol, maxit, evals,
                               vecs.ctypes.data if want_vectors else None, C.byref(nmv), C.byref(nrs))
    return dict(nconv=nconv, evals=evals, vecs=vecs, nmatvec=nmv.value, nrestart=nrs.value)


def dense_sym_eig(a):
    a = np.array(a, dtype=np.float64, order="F").copy(order="F")
    n = a.shape[0]
    ev = np.empty(n)
    flat = np.ascontiguousarray(a.T).reshape(-1)  # column-major storage
    lib().bho_dense_sym_eig(n, flat, ev, None)
    return ev


def gap_ratios(evals):
    evals = np.ascontiguousarray(evals, dtype=np.float64)
    out = np.empty(len(evals) - 2)
    lib().bho_gap_ratios(evals, len(evals), out)
    return out


def spdm(m, tags, bas, phi0, ncols=20):
    rho = np.empty((m, m))
    lib().bho_spdm(m, len(tags), tags, bas, np.ascontiguousarray(phi0), ncols, rho)
    return rho.T.copy()  # stored column-major; symmetric anyway


def condensate_fraction(rho):
    return lib().bho_condensate_fraction(rho.shape[0], np.ascontiguousarray(rho.T))


def coherence(rho):
    return lib().bho_coherence(rho.shape[0], np.ascontiguousarray(rho.T))


def point(m, tags, bas, jcsc, dU, dN, cJ, cU, cu, nb_eigen=20):
    """One grid point (src/analysis.cpp:311-337) -> dict(out3, evals, rho, nmatvec)."""
    D = len(tags)
    out3 = np.empty(3)
    ev = np.empty(nb_eigen)
    rho = np.empty((m, m))
    nmv = C.c_int(0)
    rc = lib().bho_point(m, D, tags, bas, jcsc[0], jcsc[1], jcsc[2], dU, dN, cJ, cU, cu, nb_eigen, out3,
                         ev.ctypes.data, rho.ctypes.data, C.byref(nmv))
    if rc != 0:
        raise RuntimeError("Eigenvalue computation failed.")
    return dict(out3=out3, evals=ev, rho=rho.T.copy(), nmatvec=nmv.value)
