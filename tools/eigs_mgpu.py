#!/usr/bin/env python3
"""Row-partitioned eigensolve over several GPUs (BASELINE.json config 5): one process per GPU under torchrun.
The NCCL id is created by rank 0 inside the library and broadcast with torch.distributed; everything on the data
path (all-gather of the Lanczos vector, all-reduce of the recurrence scalars) is issued by the library itself.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      tools/eigs_mgpu.py -m 14 -n 14 -U 4 --nev 2 --ncv 12 [--check] [--point]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("-m", type=int, required=True); ap.add_argument("-n", type=int, required=True)
ap.add_argument("-J", type=float, default=1.0); ap.add_argument("-U", type=float, default=4.0); ap.add_argument("-u", type=float, default=1.0)
ap.add_argument("--nev", type=int, default=20); ap.add_argument("--ncv", type=int, default=0)
ap.add_argument("--check", action="store_true", help="rank 0 also solves on one GPU and compares")
ap.add_argument("--point", action="store_true", help="full grid point (eigensolve + observables) instead of eigenvalues only")
a = ap.parse_args()
pkg = g.load_package(); capi = pkg.capi
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.tensor(list(pkg.Context.dist_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
ctx = pkg.Context(local)
ctx.dist_init(world, rank, bytes(idt.cpu().numpy().tolist()))
t0 = time.time()
ctx.setup_partitioned(a.m, a.n)
row0, nrows, _ = ctx.partition()
ncv = a.ncv or 2 * a.nev + 1
if not a.point:  # warm-up: workspace allocation, NCCL channel set-up
    ctx.eigs(a.J, a.U, a.u, nev=a.nev, ncv=ncv, kernel=capi.HV_MATRIX_FREE, order=capi.LEX, maxit=2, allow_noconv=True)
torch.cuda.synchronize(); dist.barrier()
t1 = time.time()
if a.point:
    r = ctx.point(a.J, a.U, a.u, nb_eigen=a.nev, kernel=capi.HV_MATRIX_FREE)
else:
    r = ctx.eigs(a.J, a.U, a.u, nev=a.nev, ncv=ncv, kernel=capi.HV_MATRIX_FREE, order=capi.LEX)
torch.cuda.synchronize(); dist.barrier()
t2 = time.time()
out = {"world": world, "m": a.m, "n": a.n, "D": ctx.D, "rows_rank0": nrows, "nev": a.nev, "ncv": ncv, "setup_s": t1 - t0,
       "solve_s": t2 - t1, "nmatvec": r["nmatvec"], "nrestart": r["nrestart"], "E0": float(r["evals"][0]), "E1": float(r["evals"][1])}
ok = True
if a.check and rank == 0:
    one = pkg.Context(local).setup(a.m, a.n)
    if a.point:
        s = one.point(a.J, a.U, a.u, nb_eigen=a.nev, kernel=capi.HV_MATRIX_FREE)
        ok = bool(np.allclose(s["out3"], r["out3"], rtol=1e-9, atol=1e-12) and np.abs(s["rho"] - r["rho"]).max() < 1e-10)
        out["out3"] = [float(v) for v in r["out3"]]
    else:
        one.eigs(a.J, a.U, a.u, nev=a.nev, ncv=ncv, kernel=capi.HV_MATRIX_FREE, order=capi.LEX, maxit=2, allow_noconv=True)
        s = one.eigs(a.J, a.U, a.u, nev=a.nev, ncv=ncv, kernel=capi.HV_MATRIX_FREE, order=capi.LEX)
    scale = np.maximum(np.abs(s["evals"]), abs(s["evals"][0]))
    ok = ok and bool(np.all(np.abs(s["evals"] - r["evals"]) <= 1e-10 * scale))
    out["single_gpu_solve_s"] = s["seconds"]; out["single_gpu_nmatvec"] = s["nmatvec"]
    out["max_abs_diff"] = float(np.abs(s["evals"] - r["evals"]).max()); out["check"] = "ok" if ok else "MISMATCH"
    one.close()
if rank == 0:
    print(json.dumps(out), flush=True)
ctx.dist_finalize(); ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
