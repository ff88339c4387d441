"""Driver for ncu captures (GPU box): a few launches of the H.v kernels at a given size.
usage: run_hv.py m n reps [stored|free|both] [chain|rect:LX:LY]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
pkg = g.load_package(); capi = pkg.capi
m, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
which = sys.argv[4] if len(sys.argv) > 4 else "both"
lat = sys.argv[5] if len(sys.argv) > 5 else "chain"
nbr = None
if lat.startswith("rect:"):
    lx, ly = [int(v) for v in lat.split(":")[1:]]
    nbr = capi.neighbours_rect(lx, ly)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = pkg.Context(0); ctx.set_stream(stream.cuda_stream); ctx.setup(m, n, nbr)
D = ctx.D
x = torch.empty(D, dtype=torch.float64, device="cuda"); y = torch.empty(D, dtype=torch.float64, device="cuda")
ctx.lcg_fill_dev(x.data_ptr(), D)
ref = None
for name, kid in (("stored", capi.HV_STORED), ("free", capi.HV_MATRIX_FREE)):
    if which not in (name, "both"): continue
    for _ in range(3): ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), kid)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), kid)
    b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    ab = ctx.hv_algorithmic_bytes(kid)
    if ref is None: ref = y.clone()
    err = float((y - ref).abs().max() / ref.abs().max())
    print(f"{name}: maxrel vs first {err:.1e} D={D} nnzH={ctx.hamiltonian_nnz()} {ms*1e3:.1f} us/launch  algorithmic {ab/1e6:.1f} MB -> {ab/ms/1e6:.1f} GB/s", flush=True)
