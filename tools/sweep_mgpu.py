#!/usr/bin/env python3
"""Multi-process sweep driver: one process per GPU (torchrun), grid points sharded over ranks, one all_gather
of the results at the end; rank 0 writes phase.txt.  Same options as the reference CLI.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/sweep_mgpu.py -m 12 -n 12 -J 1 -U 0 -u 0 -r 31 -s 1 -f J
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("-m", type=int, required=True); ap.add_argument("-n", type=int, required=True)
ap.add_argument("-J", type=float, default=0); ap.add_argument("-U", type=float, default=0); ap.add_argument("-u", type=float, default=0)
ap.add_argument("-r", type=float, required=True); ap.add_argument("-s", type=float, required=True)
ap.add_argument("-f", default="J"); ap.add_argument("-o", default="phase.txt"); ap.add_argument("-k", default="free")
ap.add_argument("--batch", type=int, default=4, help="grid points solved in lockstep per GPU (bh_ctx_set_batch)")
a = ap.parse_args()
pkg = g.load_package()
from bose_hubbard_phase_transition_b200 import sweep

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pkg.Context(local).setup(a.m, a.n)
ctx.set_batch(a.batch)
kern = pkg.capi.HV_STORED if a.k == "stored" else pkg.capi.HV_MATRIX_FREE
grid = sweep.make_grid(a.f, a.J, a.U, a.u, a.r, a.s)
t0 = time.time()
rows = sweep.run_sweep(lambda cJ, cU, cmu, nb: ctx.point(cJ, cU, cmu, nb, kern)["out3"], grid, world, rank, dist if world > 1 else None,
                       points_fn=lambda cJ, cU, cmu, nb: ctx.points(cJ, cU, cmu, nb, kern)[0])
if rank == 0:
    sweep.write_phase(a.o, grid, rows)
    print(f"{len(rows)} points on {world} GPU(s) in {time.time() - t0:.2f} s -> {a.o}")
if world > 1:
    dist.destroy_process_group()
