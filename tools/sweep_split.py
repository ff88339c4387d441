"""GPU box: time the matrix-free H.v variants (per-row chain kernel vs the split kernel for several cut positions and
group sizes) and check them against each other.  usage: sweep_split.py m n reps [p,p,...] [G,G,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
pkg = g.load_package(); capi = pkg.capi
m, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
# configurations "p:G:UJ:NX" (cut position, prefixes per warp, hops in flight, warps along the suffix direction)
configs = sys.argv[4].split(",") if len(sys.argv) > 4 else [f"{m // 2}:4:2:8"]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
x = y = ref = None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(label, env, cold=False):
    global x, y, ref
    for k in ("BH_FREE_VARIANT", "BH_SPLIT_P", "BH_SPLIT_G", "BH_SPLIT_UJ", "BH_SPLIT_NX"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = pkg.Context(0); ctx.set_stream(stream.cuda_stream); ctx.setup(m, n)
    D = ctx.D
    if x is None:
        x = torch.empty(D, dtype=torch.float64, device="cuda"); y = torch.empty(D, dtype=torch.float64, device="cuda")
        ctx.lcg_fill_dev(x.data_ptr(), D)
    for _ in range(3): ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), capi.HV_MATRIX_FREE)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), capi.HV_MATRIX_FREE)
    b.record(stream); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    cold_us = []
    for _ in range(5):  # L2 flushed before the launch
        with torch.cuda.stream(stream):
            flush.zero_()
        a.record(stream)
        ctx.hv_dev(1.0, 4.0, 1.0, x.data_ptr(), y.data_ptr(), capi.HV_MATRIX_FREE)
        b.record(stream); torch.cuda.synchronize()
        cold_us.append(a.elapsed_time(b) * 1e3)
    if ref is None: ref = y.clone()
    err = float((y - ref).abs().max() / ref.abs().max())
    ab = 16 * D
    print(f"m={m} n={n} {label}: {us:.1f} us warm ({ab / us / 1e3:.0f} GB/s of 16D), {min(cold_us):.1f} us L2-flushed; maxrel vs chain {err:.1e}", flush=True)
    ctx.close()


run("chain kernel", {"BH_FREE_VARIANT": "1"})
for c in configs:
    p, G, UJ, NX = c.split(":")
    run(f"split p={p} G={G} UJ={UJ} NX={NX}", {"BH_FREE_VARIANT": "2", "BH_SPLIT_P": p, "BH_SPLIT_G": G, "BH_SPLIT_UJ": UJ, "BH_SPLIT_NX": NX})
