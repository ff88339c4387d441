"""ctypes binding of libbh_b200.so (the C ABI declared in include/bh_b200.h).

This is harness plumbing for the tests and bench.py: the product is the shared library and the C++
host layer; Python only carries host buffers (numpy) or device pointers (torch) across the C ABI.
There is no fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BH_B200_LIB", os.path.join(HERE, "libbh_b200.so"))  # override: kernel-variant experiments

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NOCONV, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
LEX, TAG_SORTED, REF_SCATTER = 0, 1, 2
TERM_J, TERM_U, TERM_MU = 0, 1, 2
HV_STORED, HV_MATRIX_FREE, HV_USER, HV_HYBRID = 0, 1, 2, 3
PROF_CLASSES = ["hv_free", "hv_batch2", "hv_batch4", "hv_stored", "step", "restart", "gram", "spdm", "small"]

# every symbol include/bh_b200.h declares (checked by tests/test_abi.py against the header)
SYMBOLS = [
    "bh_ctx_create", "bh_ctx_destroy", "bh_last_error", "bh_ctx_set_stream", "bh_ctx_launch_count", "bh_ctx_transfer_bytes",
    "bh_neighbours_chain", "bh_neighbours_rect", "bh_dimension", "bh_setup", "bh_basis", "bh_rank",
    "bh_term_nnz", "bh_term_csc", "bh_hamiltonian_nnz", "bh_hamiltonian_csc", "bh_hv", "bh_hv_dev", "bh_eigs",
    "bh_spdm", "bh_gap_ratios", "bh_condensate_fraction", "bh_coherence", "bh_point", "bh_points",
    "bh_lcg_fill_dev", "bh_hv_algorithmic_bytes", "bh_load_matrix", "bh_ctx_set_batch",
    "bh_ctx_profile_enable", "bh_ctx_profile_read", "bh_host_register", "bh_host_unregister", "bh_thermal_weights", "bh_density_matrix",
    "bh_dist_unique_id", "bh_dist_init", "bh_dist_finalize", "bh_setup_partitioned", "bh_partition",
]


class EigsInfo(C.Structure):
    _fields_ = [("nconv", C.c_int32), ("nmatvec", C.c_int32), ("nrestart", C.c_int32), ("nreorth", C.c_int32),
                ("seconds", C.c_double)]


class BhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bh_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libbh_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    L.bh_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.bh_ctx_destroy.argtypes = [vp]
    L.bh_last_error.argtypes = [vp]
    L.bh_last_error.restype = C.c_char_p
    L.bh_ctx_set_stream.argtypes = [vp, vp]
    L.bh_ctx_launch_count.argtypes = [vp]
    L.bh_ctx_launch_count.restype = C.c_int64
    L.bh_ctx_transfer_bytes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.bh_neighbours_chain.argtypes = [C.c_int, C.c_int, vp, vp]
    L.bh_neighbours_rect.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.bh_dimension.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.bh_setup.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.bh_basis.argtypes = [vp, C.c_int, vp, vp]
    L.bh_rank.argtypes = [vp, C.c_int, vp, C.c_int64, vp]
    L.bh_term_nnz.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.bh_term_csc.argtypes = [vp, C.c_int, C.c_double, C.c_int, vp, vp, vp]
    L.bh_hamiltonian_nnz.argtypes = [vp, C.POINTER(C.c_int64)]
    L.bh_hamiltonian_csc.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, vp, vp, vp]
    L.bh_hv.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp]
    L.bh_hv_dev.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, vp, vp]
    L.bh_eigs.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                          C.c_int, vp, vp, C.POINTER(EigsInfo)]
    L.bh_spdm.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    L.bh_gap_ratios.argtypes = [vp, C.c_int, vp]
    L.bh_condensate_fraction.argtypes = [C.c_int, vp, dp]
    L.bh_coherence.argtypes = [C.c_int, vp, dp]
    L.bh_point.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp, vp, C.POINTER(EigsInfo)]
    L.bh_points.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp, vp]
    L.bh_ctx_set_batch.argtypes = [vp, C.c_int]
    L.bh_thermal_weights.argtypes = [vp, C.c_int, C.c_double, vp]
    L.bh_density_matrix.argtypes = [vp, C.c_int64, C.c_int, vp, vp, C.c_double, vp]
    L.bh_host_register.argtypes = [vp, C.c_int64]
    L.bh_host_unregister.argtypes = [vp]
    L.bh_ctx_profile_enable.argtypes = [vp, C.c_int]
    L.bh_ctx_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), dp, dp]
    L.bh_dist_unique_id.argtypes = [vp]
    L.bh_dist_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.bh_dist_finalize.argtypes = [vp]
    L.bh_setup_partitioned.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.bh_partition.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.bh_load_matrix.argtypes = [vp, C.c_int64, vp, vp, vp]
    L.bh_lcg_fill_dev.argtypes = [vp, vp, C.c_int64]
    L.bh_hv_algorithmic_bytes.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def thermal_weights(evals, temperature):
    e = np.ascontiguousarray(evals, dtype=np.float64)
    w = np.empty(len(e))
    rc = load().bh_thermal_weights(_ptr(e), len(e), float(temperature), _ptr(w))
    if rc:
        raise BhError(rc, "bh_thermal_weights: bad argument")
    return w


def host_register(a):
    """Page-lock a numpy array owned by the caller (bh_host_register); returns True on success."""
    return load().bh_host_register(_ptr(a), a.nbytes) == OK


def host_unregister(a):
    return load().bh_host_unregister(_ptr(a)) == OK


def dimension(m, n):
    d = C.c_int64(0)
    rc = load().bh_dimension(m, n, C.byref(d))
    if rc:
        raise BhError(rc, "bh_dimension: bad argument")
    return d.value


def neighbours_chain(m, closed=True):
    L = load()
    ptr = np.zeros(m + 1, dtype=np.int32)
    L.bh_neighbours_chain(m, int(closed), _ptr(ptr), None)
    idx = np.zeros(max(int(ptr[m]), 1), dtype=np.int32)
    L.bh_neighbours_chain(m, int(closed), _ptr(ptr), _ptr(idx))
    return ptr, idx[: ptr[m]].copy()


def neighbours_rect(lx, ly=1, lz=1, closed=True):
    L = load()
    m = lx * ly * lz
    ptr = np.zeros(m + 1, dtype=np.int32)
    L.bh_neighbours_rect(lx, ly, lz, int(closed), _ptr(ptr), None)
    idx = np.zeros(max(int(ptr[m]), 1), dtype=np.int32)
    L.bh_neighbours_rect(lx, ly, lz, int(closed), _ptr(ptr), _ptr(idx))
    return ptr, idx[: ptr[m]].copy()


def gap_ratios(evals):
    e = np.ascontiguousarray(evals, dtype=np.float64)
    out = np.empty(len(e) - 2)
    rc = load().bh_gap_ratios(_ptr(e), len(e), _ptr(out))
    if rc:
        raise BhError(rc, "bh_gap_ratios")
    return out


def condensate_fraction(rho):
    r = np.ascontiguousarray(np.asarray(rho, dtype=np.float64).T)
    out = C.c_double(0)
    load().bh_condensate_fraction(r.shape[0], _ptr(r), C.byref(out))
    return out.value


def coherence(rho):
    r = np.ascontiguousarray(np.asarray(rho, dtype=np.float64).T)
    out = C.c_double(0)
    load().bh_coherence(r.shape[0], _ptr(r), C.byref(out))
    return out.value


class Context:
    """One GPU + one system (m sites, n bosons, neighbour list)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.bh_ctx_create(device, C.byref(h))
        if rc:
            raise BhError(rc, self.L.bh_last_error(None).decode())
        self.h = h
        self.m = self.n = self.D = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.bh_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise BhError(rc, self.L.bh_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self.L.bh_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_batch(self, batch):
        """bh_points solves `batch` (1..4) grid points in lockstep, sharing their H.v launches."""
        self._check(self.L.bh_ctx_set_batch(self.h, int(batch)))
        return self

    def profile_enable(self, on=True):
        """Bracket every launch of the main kernel classes with CUDA events on the launching stream (bench.py's roofline)."""
        self._check(self.L.bh_ctx_profile_enable(self.h, int(on)))

    def profile_read(self):
        """-> {class: dict(launches, ms, bytes)} since the last read."""
        out = {}
        for cls, name in enumerate(PROF_CLASSES):
            n, ms, by = C.c_int64(0), C.c_double(0), C.c_double(0)
            self._check(self.L.bh_ctx_profile_read(self.h, cls, C.byref(n), C.byref(ms), C.byref(by)))
            out[name] = dict(launches=n.value, ms=ms.value, bytes=by.value)
        return out

    def launch_count(self):
        return self.L.bh_ctx_launch_count(self.h)

    def transfer_bytes(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self.L.bh_ctx_transfer_bytes(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def setup(self, m, n, nbr=None):
        ptr, idx = nbr if nbr is not None else neighbours_chain(m)
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._check(self.L.bh_setup(self.h, m, n, _ptr(ptr), _ptr(idx)))
        self.m, self.n, self.D = m, n, dimension(m, n)
        return self

    # ---- row-partitioned solve over several GPUs (one process per GPU) ----
    @staticmethod
    def dist_unique_id():
        buf = (C.c_ubyte * 128)()
        rc = load().bh_dist_unique_id(buf)
        if rc:
            raise BhError(rc, load().bh_last_error(None).decode())
        return bytes(buf)

    def dist_init(self, world, rank, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self.L.bh_dist_init(self.h, world, rank, buf))

    def dist_finalize(self):
        self._check(self.L.bh_dist_finalize(self.h))

    def setup_partitioned(self, m, n, nbr=None):
        ptr, idx = nbr if nbr is not None else neighbours_chain(m)
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._check(self.L.bh_setup_partitioned(self.h, m, n, _ptr(ptr), _ptr(idx)))
        self.m, self.n, self.D = m, n, dimension(m, n)
        return self

    def partition(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self.L.bh_partition(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def load_matrix(self, outer, inner, val):
        outer = np.ascontiguousarray(outer, dtype=np.int32)
        inner = np.ascontiguousarray(inner, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        D = len(outer) - 1
        self._check(self.L.bh_load_matrix(self.h, D, _ptr(outer), _ptr(inner), _ptr(val)))
        self.m = self.n = 0
        self.D = D
        return self

    def basis(self, order=TAG_SORTED):
        tags = np.empty(self.D)
        bas = np.empty((self.D, self.m))
        self._check(self.L.bh_basis(self.h, order, _ptr(tags), _ptr(bas)))
        return tags, bas

    def rank(self, states, order=TAG_SORTED):
        st = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, self.m)
        out = np.empty(st.shape[0], dtype=np.int32)
        self._check(self.L.bh_rank(self.h, order, _ptr(st), st.shape[0], _ptr(out)))
        return out

    def term_nnz(self, term):
        v = C.c_int64(0)
        self._check(self.L.bh_term_nnz(self.h, term, C.byref(v)))
        return v.value

    def hamiltonian_nnz(self):
        v = C.c_int64(0)
        self._check(self.L.bh_hamiltonian_nnz(self.h, C.byref(v)))
        return v.value

    def term_csc(self, term, coef=1.0, order=TAG_SORTED):
        nnz = self.term_nnz(term)
        outer = np.empty(self.D + 1, dtype=np.int32)
        inner = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz)
        self._check(self.L.bh_term_csc(self.h, term, coef, order, _ptr(outer), _ptr(inner), _ptr(val)))
        return outer, inner, val

    def hamiltonian_csc(self, cJ, cU, cmu, order=TAG_SORTED):
        nnz = self.hamiltonian_nnz()
        outer = np.empty(self.D + 1, dtype=np.int32)
        inner = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz)
        self._check(self.L.bh_hamiltonian_csc(self.h, cJ, cU, cmu, order, _ptr(outer), _ptr(inner), _ptr(val)))
        return outer, inner, val

    def hv(self, cJ, cU, cmu, x, kernel=HV_STORED, order=TAG_SORTED, out=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = out if out is not None else np.empty(self.D)
        self._check(self.L.bh_hv(self.h, cJ, cU, cmu, kernel, order, _ptr(x), _ptr(y)))
        return y

    def hv_dev(self, cJ, cU, cmu, x_ptr, y_ptr, kernel=HV_STORED):
        self._check(self.L.bh_hv_dev(self.h, cJ, cU, cmu, kernel, C.c_void_p(x_ptr), C.c_void_p(y_ptr)))

    def lcg_fill_dev(self, x_ptr, count):
        self._check(self.L.bh_lcg_fill_dev(self.h, C.c_void_p(x_ptr), count))

    def hv_algorithmic_bytes(self, kernel=HV_STORED):
        v = C.c_int64(0)
        self._check(self.L.bh_hv_algorithmic_bytes(self.h, kernel, C.byref(v)))
        return v.value

    def eigs(self, cJ, cU, cmu, nev=20, ncv=None, tol=1e-10, maxit=1000, kernel=HV_STORED, order=TAG_SORTED,
             want_vectors=False, allow_noconv=False):
        ncv = ncv or 2 * nev + 1
        ev = np.full(nev, np.nan)
        nrows = self.partition()[1] if want_vectors else 0
        vecs = np.empty((nev, nrows)) if want_vectors else None
        info = EigsInfo()
        rc = self.L.bh_eigs(self.h, cJ, cU, cmu, nev, ncv, tol, maxit, kernel, order, _ptr(ev), _ptr(vecs),
                            C.byref(info))
        if rc and not (allow_noconv and rc == ERR_NOCONV):
            self._check(rc)
        return dict(evals=ev, vecs=vecs, nconv=info.nconv, nmatvec=info.nmatvec, nrestart=info.nrestart,
                    nreorth=info.nreorth, seconds=info.seconds, rc=rc)

    def density_matrix(self, evals, vecs, temperature):
        """rho = sum_k w_k u_k u_k^T (D x D) from eigenvectors given as rows of `vecs` (what eigs(want_vectors=True) returns)."""
        e = np.ascontiguousarray(evals, dtype=np.float64)
        u = np.ascontiguousarray(vecs, dtype=np.float64)   # [k][i] = column-major D x nev
        D = u.shape[1]
        rho = np.empty((D, D))
        self._check(self.L.bh_density_matrix(self.h, D, len(e), _ptr(e), _ptr(u), float(temperature), _ptr(rho)))
        return rho.T.copy()

    def spdm(self, phi, ncols=20, order=TAG_SORTED):
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        rho = np.empty((self.m, self.m))
        self._check(self.L.bh_spdm(self.h, order, _ptr(phi), ncols, _ptr(rho)))
        return rho.T.copy()

    def point(self, cJ, cU, cmu, nb_eigen=20, kernel=HV_STORED):
        out3 = np.empty(3)
        ev = np.empty(nb_eigen)
        rho = np.empty((self.m, self.m))
        info = EigsInfo()
        self._check(self.L.bh_point(self.h, cJ, cU, cmu, nb_eigen, kernel, _ptr(out3), _ptr(ev), _ptr(rho),
                                    C.byref(info)))
        return dict(out3=out3, evals=ev, rho=rho.T.copy(), nmatvec=info.nmatvec, nrestart=info.nrestart,
                    nreorth=info.nreorth, seconds=info.seconds)

    def points(self, cJ, cU, cmu, nb_eigen=20, kernel=HV_STORED):
        cJ = np.ascontiguousarray(cJ, dtype=np.float64)
        cU = np.ascontiguousarray(cU, dtype=np.float64)
        cmu = np.ascontiguousarray(cmu, dtype=np.float64)
        npts = len(cJ)
        out3 = np.empty((npts, 3))
        infos = (EigsInfo * max(npts, 1))()
        self._check(self.L.bh_points(self.h, _ptr(cJ), _ptr(cU), _ptr(cmu), npts, nb_eigen, kernel, _ptr(out3), infos))
        return out3, [dict(nmatvec=i.nmatvec, nrestart=i.nrestart, seconds=i.seconds) for i in infos[:npts]]
