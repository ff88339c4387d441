#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/s_b8_check.log 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0).setup(12, 12)
U = np.array([3.0, 9.0, 17.0, 26.0, 1.0, 5.0, 12.0, 31.0, 7.0, 22.0])
mu = np.arange(10.0)
ref = None
for b in (4, 8):
    ctx.set_batch(b)
    o3, infos = ctx.points(np.ones(10), U, mu, kernel=pkg.capi.HV_MATRIX_FREE)
    if ref is None: ref = (o3.copy(), [i["nmatvec"] for i in infos])
    print("batch", b, "bit-identical to batch 4:", bool((o3 == ref[0]).all()), [i["nmatvec"] for i in infos] == ref[1], flush=True)
PY
cat gpurun_out/s_b8_check.log
for B in 4 8; do for P in 16 32; do
( timeout 600 python bench.py --steps 3 --warmup 2 --batch $B --points-per-step $P --no-c5 --no-stored --no-small --no-cpu-baseline 2> gpurun_out/s_bench_b${B}_p$P.err ) > gpurun_out/s_bench_b${B}_p$P.json
done; done
python - <<'PY'
import json
for B in (4, 8):
    for P in (16, 32):
        try:
            d = json.loads([l for l in open(f"gpurun_out/s_bench_b{B}_p{P}.json") if l.startswith("{")][-1])
            print(B, P, round(d["value"], 3), round(d["e2e"]["value"], 3), [(e["class"], round(e["share_of_step"] or 0, 3), round(e["ms_per_launch"], 4)) for e in d["roofline_path"][:6]])
        except Exception as ex:
            print(B, P, "failed", ex)
PY
tail -2 gpurun_out/s_bench_b8_p16.err
