"""Runner for the real (patched) reference binaries in oracle/_ref -- TEST INFRASTRUCTURE ONLY.

oracle/_ref/ is built by `make -C oracle ref` where /root/reference exists (this
container); the binaries travel to the GPU box but /root/reference does not.
"""
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HARNESS = os.path.join(REF_DIR, "ref_harness")
HARNESS_UNPATCHED = os.path.join(REF_DIR, "ref_harness_unpatched")
CLI = os.path.join(REF_DIR, "QuantumProject")


def available():
    return os.path.exists(HARNESS) and os.access(HARNESS, os.X_OK)


def _run(exe, args, names, env=None, timeout=3600):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "o")
        argv = [exe] + [str(a) if a != "@out" else out for a in args]
        e = dict(os.environ)
        if env:
            e.update(env)
        p = subprocess.run(argv, capture_output=True, text=True, env=e, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError(f"{argv}: rc={p.returncode}\n{p.stdout}\n{p.stderr}")
        info = json.loads(p.stdout.strip().splitlines()[-1])
        res = {}
        for name, dt in names:
            ext = "f64" if dt == np.float64 else "i32"
            f = f"{out}.{name}.{ext}"
            if os.path.exists(f):
                res[name] = np.fromfile(f, dtype=dt)
        return info, res


def basis(m, n, unpatched=False):
    info, r = _run(HARNESS_UNPATCHED if unpatched else HARNESS, ["basis", m, n, "@out"],
                   [("tags", np.float64), ("basis", np.float64)])
    return r["tags"], r["basis"].reshape(-1, m), info


def csc(m, n, term, lattice="chain"):
    info, r = _run(HARNESS, ["csc", m, n, term, lattice, "@out"],
                   [("outer", np.int32), ("inner", np.int32), ("val", np.float64)])
    return (r["outer"], r["inner"], r["val"]), info


def hsum(m, n, cJ, cU, cu, lattice="chain"):
    info, r = _run(HARNESS, ["hsum", m, n, cJ, cU, cu, lattice, "@out"],
                   [("outer", np.int32), ("inner", np.int32), ("val", np.float64)])
    return (r["outer"], r["inner"], r["val"]), info


def eigs(m, n, cJ, cU, cu, nev=20, lattice="chain"):
    info, r = _run(HARNESS, ["eigs", m, n, cJ, cU, cu, nev, lattice, "@out"],
                   [("evals", np.float64), ("rho", np.float64), ("out5", np.float64)])
    r["rho"] = r["rho"].reshape(m, m).T
    return r, info


def hv(m, n, cJ, cU, cu, reps=1, lattice="chain", want=True):
    info, r = _run(HARNESS, ["hv", m, n, cJ, cU, cu, reps, lattice, "@out" if want else "-"],
                   [("x", np.float64), ("y", np.float64)])
    return r, info


def points(m, n, fixed, cfix, p1min, p2min, step, n1, n2, threads=0, lattice="chain", want=True, timeout=7200):
    info, r = _run(HARNESS, ["points", m, n, fixed, cfix, p1min, p2min, step, n1, n2, threads, lattice,
                             "@out" if want else "-"], [("out5", np.float64), ("evals", np.float64)], timeout=timeout)
    if "out5" in r:
        r["out5"] = r["out5"].reshape(-1, 5)
        r["evals"] = r["evals"].reshape(-1, 20)
    return r, info


def thermal(m, n, cJ, cU, cu, T):
    info, r = _run(HARNESS, ["thermal", m, n, cJ, cU, cu, T, "@out"], [("dm", np.float64), ("evals", np.float64), ("out2", np.float64)])
    D = info["D"]
    r["dm"] = r["dm"].reshape(D, D).T
    return r


def max_basis(m, n):
    info, r = _run(HARNESS, ["maxbasis", m, n, "@out"], [("tags", np.float64), ("basis", np.float64)])
    return r["tags"], r["basis"].reshape(-1, m)


def max_hamiltonian(m, nmin, nmax, J, U, mu, lattice="chain"):
    info, r = _run(HARNESS, ["maxham", m, nmin, nmax, J, U, mu, lattice, "@out"],
                   [("outer", np.int32), ("inner", np.int32), ("val", np.float64)])
    return r["outer"], r["inner"], r["val"]


def partial(m, n, cJ, cU, cu, maxit, threads=0, lattice="chain", timeout=7200):
    info, _ = _run(HARNESS, ["partial", m, n, cJ, cU, cu, maxit, threads, lattice], [], timeout=timeout)
    return info


def cli_phase(args, threads=None, timeout=7200):
    """Run the patched reference CLI in a scratch dir; returns the text of phase.txt."""
    with tempfile.TemporaryDirectory() as td:
        e = dict(os.environ)
        if threads:
            e["OMP_NUM_THREADS"] = str(threads)
        p = subprocess.run([CLI] + [str(a) for a in args], cwd=td, capture_output=True, text=True, env=e,
                           timeout=timeout)
        f = os.path.join(td, "phase.txt")
        if not os.path.exists(f):
            raise RuntimeError(f"no phase.txt: rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
        return open(f).read()
