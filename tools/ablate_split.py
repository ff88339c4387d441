"""GPU box: ablation timings of the split H.v kernel (which part of the row costs what).  usage: ablate_split.py m n reps G UJ"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
pkg = g.load_package(); capi = pkg.capi
m, n, reps, G, UJ = [int(v) for v in sys.argv[1:6]]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
x = y = None
for ab in (0, 1, 2, 4, 3, 5, 6, 7):
    os.environ.update({"BH_FREE_VARIANT": "2", "BH_SPLIT_G": str(G), "BH_SPLIT_UJ": str(UJ), "BH_SPLIT_ABLATE": str(ab)})
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
    skipped = [name for bit, name in ((1, "suffix"), (2, "prefix"), (4, "cross")) if ab & bit]
    print(f"m={m} n={n} G={G} UJ={UJ} skip {'+'.join(skipped) or 'nothing'}: {a.elapsed_time(b) / reps * 1e3:.1f} us", flush=True)
    ctx.close()
