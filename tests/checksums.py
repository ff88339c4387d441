"""Exact integer checksums used by the golden fixtures (tests/golden/make_golden.py) and the tests."""
import numpy as np


def checksum_basis(t, b):
    """Exact integer checksums (independent of BLAS summation order): wrap-around uint64 sum of the tag bit
    patterns weighted by position, their xor, and an int64 position-weighted sum of the occupations."""
    tb = np.ascontiguousarray(t, dtype=np.float64).view(np.uint64)
    w = np.arange(1, len(t) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        s1 = np.uint64((tb * w).sum(dtype=np.uint64))
    x1 = np.bitwise_xor.reduce(tb)
    occ = (b.astype(np.int64) * (np.arange(b.shape[1], dtype=np.int64) + 1)).sum(1)
    s2 = np.int64((occ * w.astype(np.int64)).sum())
    return np.array([np.uint64(len(t)), s1, x1, np.uint64(s2)], dtype=np.uint64)


def checksum_csc(o, i, v):
    vb = np.ascontiguousarray(v, dtype=np.float64).view(np.uint64)
    w = np.arange(1, len(i) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return np.array([np.uint64(len(i)), np.uint64((vb * w).sum(dtype=np.uint64)), np.bitwise_xor.reduce(vb),
                         np.uint64((i.astype(np.uint64) * w).sum(dtype=np.uint64)),
                         np.uint64((o.astype(np.uint64) * np.arange(1, len(o) + 1, dtype=np.uint64)).sum(dtype=np.uint64))],
                        dtype=np.uint64)
