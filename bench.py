#!/usr/bin/env python3
"""bench.py -- phase-diagram points/s and H.v GB/s vs HBM peak at m = n = 12 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config 3 of BASELINE.json, SURVEY.md section 8d "C3"): closed chain m = 12 sites, n = 12 bosons
(D = 1 352 078, nnz(H) = 18 282 446), the 32 x 32 grid of `-J 1 -U 0 -u 0 -r 31 -s 1 -f J`
(J-coefficient 1, U-coefficient 1..32, mu 0..31).  A step = one list of P (default 8) grid points per GPU (eigensolve for
the 20 lowest levels + gap ratio + SPDM + condensate fraction + coherence; the points of a batch are solved in lockstep and
share their H.v launches, --batch, results identical point by point); every rank gets the same U values (a fixed
pseudo-random order over the grid's 32) at different mu; per-GPU work is fixed as N grows (weak scaling, no data-path
collective).

  value     points/s, whole job, basis + stored H resident in HBM, timed with CUDA events on the launching stream
  e2e       points/s through the C ABI from a cold context: bh_setup + bh_points with host buffers, wall clock,
            PCIe bytes counted by the library (bh_ctx_transfer_bytes)
  roofline  the stored-CSR H.v kernel (K3) timed alone with CUDA events: algorithmic bytes / duration vs the
            measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
            the (patched) reference's own solver call on the host cores: `nproc` concurrent copies of one C3 point,
            each stopped after a bounded number of restarts (a full C3 point takes ~4 min per core), points/s
            extrapolated by the H.v-count ratio to the converged solve
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRID = 32  # -r 31 -s 1  ->  int(31 / 1) + 1 points per axis
# H.v applications of the REFERENCE algorithm (Spectra, nev 20, ncv 41) per converged C3 point, mean over the 32 U values
# of the grid: measured with this library's plain mode (BH_CHEB_DEGREE=1), whose counts track Spectra's within a few %
# (DESIGN.md section 4); used only to scale the bounded CPU sample to a full solve.
REF_MEAN_MATVECS = 2167


def grid_points(m_unused=None):
    """(cJ, cU, cmu) of the C3 grid in the reference's loop order (src/analysis.cpp:303-308, -f J mode)."""
    p1 = 1.0 + np.arange(GRID) * 1.0  # U coefficient sweeps [J, J + r]  
    p2 = 0.0 + np.arange(GRID) * 1.0  # mu
    cU, cmu = np.meshgrid(p1, p2, indexing="ij")
    return np.ones(GRID * GRID), cU.reshape(-1), cmu.reshape(-1)


def peaks():
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        return json.load(open(f)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_sample(maxit, threads, full_matvecs):
    """Bounded CPU sample of one C3 point (J=1, U=4, mu=1) on `threads` cores with oracle/_ref; returns the dict
    for cpu_baseline.  full_matvecs = H.v count of the converged solve (from the GPU solver, same algorithm)."""
    import ref_lib as R
    if not R.available():
        return None
    info = R.partial(12, 12, 1, 4, 1, maxit, threads)
    per_point = info["seconds"] * (full_matvecs / max(info["nmatvec"], 1))
    return {
        "value": info["threads"] / per_point, "unit": "points/s", "cores": info["threads"], "kind": "reference",
        "sample": (f"{info['threads']} concurrent copies (one per core) of the C3 point (J=1,U=4,mu=1) through the patched "
                   f"reference solver call (Spectra GenEigsSolver nev=20 ncv=41), stopped after {info['nrestart']} restart(s) = "
                   f"{info['nmatvec']} H.v in {info['seconds']:.1f} s; scaled to the {full_matvecs} H.v of the converged solve"),
        "seconds": info["seconds"], "sample_matvecs": info["nmatvec"], "setup_seconds": info["setup_seconds"],
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ref_lib as R
    cfg = {"workload": "C3: closed chain m=12 n=12 (D=1352078), 32x32 grid of -J 1 -U 0 -u 0 -r 31 -s 1 -f J",
           "points_per_step": os.cpu_count()}
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (needs /root/reference at build time)"}))
        return 0
    full = args.full_matvecs
    vals = []
    for s in range(args.warmup + args.steps):
        r = reference_sample(args.ref_maxit, os.cpu_count(), full)
        if s >= args.warmup:
            vals.append(r)
    v = float(np.mean([x["value"] for x in vals]))
    ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
    line = {
        "impl": "reference", "metric": "phase_diagram_points_per_sec", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: vals[-1][k] for k in ("unit", "cores", "kind", "sample")} | {"value": v},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points-per-step", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4,
                    help="grid points solved in lockstep per GPU (bh_ctx_set_batch): their Chebyshev-filter H.v launches are shared")
    ap.add_argument("--kernel", default="free", choices=["stored", "free"],
                    help="H.v kernel of the eigensolver: matrix-free (chain-specialised, faster at m=n=12) or stored SELL-32")
    ap.add_argument("--hv-reps", type=int, default=50)
    ap.add_argument("--ref-maxit", type=int, default=1)
    ap.add_argument("--full-matvecs", type=int, default=REF_MEAN_MATVECS,
                    help="H.v count of a converged C3 solve used to scale the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the m=n=14 matrix-free H.v timing")
    ap.add_argument("--m", type=int, default=12)
    ap.add_argument("--n", type=int, default=12)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    capi = pkg.capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = world
    P = args.points_per_step
    K, W = args.steps, args.warmup
    m, n = args.m, args.n
    kernel = capi.HV_STORED if args.kernel == "stored" else capi.HV_MATRIX_FREE

    cJ, cU, cmu = grid_points()
    uperm = np.random.default_rng(0).permutation(GRID)

    def step_points(s):
        """Grid points of step s on this rank: every rank gets the same U values (fixed pseudo-random order over the 32
        of the grid) at different mu values, so the per-rank work is equal by construction -- on the real grid every U
        occurs with all 32 mu, and the cost of a point does not depend on mu (a pure shift of the spectrum)."""
        idx = []
        for q in range(P):
            iu = int(uperm[(s * P + q) % GRID])
            imu = (rank + s * P + q) % GRID
            idx.append(iu * GRID + imu)
        return cJ[idx], cU[idx], cmu[idx]

    # a dedicated (non-null) stream: the library launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    # ---- resident context: set-up outside the timed region ----
    ctx = pkg.Context(local)
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.setup(m, n)
    ctx.set_batch(args.batch)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        ctx.points(*step_points(s), kernel=kernel)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    matvecs = []
    for s in range(W, W + K):
        out3, infos = ctx.points(*step_points(s), kernel=kernel)
        matvecs += [i["nmatvec"] for i in infos]
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    dev_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = (N * P * K) / (dev_ms * 1e-3)

    # ---- e2e: cold context through the C ABI with host buffers, wall clock, set-up inside ----
    barrier()
    w0 = time.perf_counter()
    ctx2 = pkg.Context(local)
    ctx2.setup(m, n)
    ctx2.set_batch(args.batch)
    for s in range(W, W + K):
        ctx2.points(*step_points(s), kernel=kernel)
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    h2d, d2h = ctx2.transfer_bytes()
    ctx2.close()
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    e2e = {"value": (N * P * K) / wall, "unit": "points/s", "h2d_bytes_per_step": int(h2d / K),
           "d2h_bytes_per_step": int(d2h / K), "includes": "bh_setup (basis + CSR build) + bh_points, host buffers"}

    # ---- roofline: the stored-CSR H.v kernel alone (rank 0's GPU; every rank runs it to stay in step) ----
    D = ctx.D
    x = torch.empty(D, dtype=torch.float64, device="cuda")
    y = torch.empty(D, dtype=torch.float64, device="cuda")
    ctx.lcg_fill_dev(x.data_ptr(), D)
    hv = {}
    for name, kid in (("stored", capi.HV_STORED), ("matrix_free", capi.HV_MATRIX_FREE)):
        for _ in range(5):
            ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), kid)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.hv_reps):
            ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), kid)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.hv_reps
        ab = ctx.hv_algorithmic_bytes(kid)
        hv[name] = {"ms": ms, "algorithmic_bytes": ab, "gbs": ab / (ms * 1e-3) / 1e9}
    # MatOp seam with host vectors (H2D + kernel + D2H), the call Spectra would make
    xh = np.random.default_rng(1).uniform(-0.5, 0.5, D)
    ctx.hv(1.0, 4.0, 1.0, xh, kernel=capi.HV_STORED, order=capi.LEX)
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.hv(1.0, 4.0, 1.0, xh, kernel=capi.HV_STORED, order=capi.LEX)
    hv_host_ms = (time.perf_counter() - t0) / 5 * 1e3

    # config 5 shape for the record (N = 1 only): matrix-free H.v at m = n = 14, algorithmic bytes 16 D
    c5 = None
    if world == 1 and not args.no_c5:
        try:
            c14 = pkg.Context(local)
            c14.set_stream(stream.cuda_stream)
            c14.setup(14, 14)
            D14 = c14.D
            x14 = torch.empty(D14, dtype=torch.float64, device="cuda")
            y14 = torch.empty(D14, dtype=torch.float64, device="cuda")
            c14.lcg_fill_dev(x14.data_ptr(), D14)
            for _ in range(3):
                c14.hv_dev(1.0, 4.0, 1.0, x14.data_ptr(), y14.data_ptr(), capi.HV_MATRIX_FREE)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10):
                c14.hv_dev(1.0, 4.0, 1.0, x14.data_ptr(), y14.data_ptr(), capi.HV_MATRIX_FREE)
            b.record(stream)
            torch.cuda.synchronize()
            ms14 = a.elapsed_time(b) / 10
            ab14 = c14.hv_algorithmic_bytes(capi.HV_MATRIX_FREE)
            c5 = {"workload": "C5 shape: closed chain m=14 n=14 (D=20058300), matrix-free H.v", "ms": ms14,
                  "algorithmic_bytes": ab14, "gbs": ab14 / (ms14 * 1e-3) / 1e9}
            c14.close()
            del x14, y14
        except Exception as ex:
            c5 = {"error": str(ex)}

    peak, peak_src = peaks()
    if c5 and "gbs" in c5:
        c5["frac_of_hbm_peak"] = c5["gbs"] / peak
    traffic = None
    tf = os.path.join(ROOT, "profiles", "hv_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("k_hv_sell_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_hv_sell (stored H.v, K3, SELL-32 layout)", "achieved": hv["stored"]["gbs"], "peak": peak,
                "unit": "GB/s", "frac": hv["stored"]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": hv["stored"]["algorithmic_bytes"], "ms_per_launch": hv["stored"]["ms"]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    full_mv = args.full_matvecs  # the reference algorithm's count, not the accelerated solver's
    cpu = None
    if N == 1 and not args.no_cpu_baseline:
        try:
            cpu = reference_sample(args.ref_maxit, os.cpu_count(), full_mv)
        except Exception as ex:  # the baseline is a report, never a reason to lose the GPU line
            cpu = {"value": None, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        if cpu is None:
            cpu = {"value": None, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": "oracle/_ref not built on this host"}

    line = {
        "metric": "phase_diagram_points_per_sec", "value": value, "unit": "points/s", "n_gpus": N, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C3: closed chain m=12 n=12 (D=1352078, nnz(H)=18282446), 32x32 grid of -J 1 -U 0 -u 0 -r 31 -s 1 -f J",
                   "m": m, "n": n, "points_per_gpu_per_step": P, "nev": 20, "ncv": 41, "tol": 1e-10, "hv_kernel": args.kernel, "lockstep_batch": args.batch,
                   "solver": "thick-restart Lanczos on a degree-%s Chebyshev filter of H + Rayleigh-Ritz of H" % os.environ.get("BH_CHEB_DEGREE", "8"),
                   "l2": "inputs larger than L2 (Krylov basis 454 MB per point; stored H 433 MB for the roofline kernel)",
                   "mean_matvecs_per_point": int(np.mean(matvecs)) if matvecs else None, "setup_seconds": setup_s},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "hv": {"stored": hv["stored"], "matrix_free": hv["matrix_free"], "host_vectors_ms": hv_host_ms,
               "note": "roofline = the stored H.v the metric names (C3); the sweep itself runs the matrix-free chain kernel "
                       "(instruction-bound, 8x less HBM traffic), whose 16*D figure is listed here and for C5 below"},
        "c5_matrix_free_hv": c5,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
