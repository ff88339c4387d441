#!/usr/bin/env python3
"""Time the row-partitioned matrix-free H.v (config 5) under torchrun: CUDA events on the context's stream, max over ranks.
BH_DIST_ALLGATHER=1 selects the all-gather baseline; BH_HALO_ABLATE (1 no exchange, 2 no local phase, 4 no remote phase)
isolates the parts of the overlapped form (results are then wrong: timing probe only)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as g

m = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pkg = g.load_package(); capi = pkg.capi
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.tensor(list(pkg.Context.dist_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = pkg.Context(local); ctx.set_stream(stream.cuda_stream)
ctx.dist_init(world, rank, bytes(idt.cpu().numpy().tolist()))
ctx.setup_partitioned(m, m)
row0, nrows, slice_len = ctx.partition()
x = torch.zeros(slice_len, dtype=torch.float64, device="cuda"); y = torch.zeros(slice_len, dtype=torch.float64, device="cuda")
ctx.lcg_fill_dev(x.data_ptr(), nrows)
for _ in range(3):
    ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), capi.HV_MATRIX_FREE)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(reps):
    ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), capi.HV_MATRIX_FREE)
b.record(stream)
torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "m": m, "ms_per_hv": float(t.item()), "env": {k: v for k, v in os.environ.items() if k.startswith("BH_")}}), flush=True)
dist.barrier()
ctx.dist_finalize(); ctx.close()
dist.destroy_process_group()
