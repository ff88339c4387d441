#!/bin/bash
# GPU session D (1 GPU): full GPU test-suite after the solver / boundary changes, bench with and without the quick stage 1
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/d_pytest.log
( timeout 400 python bench.py --steps 3 --warmup 3 --no-c5 --no-small --no-cpu-baseline 2> gpurun_out/d_bench.err ) > gpurun_out/d_bench_quick.json
( BH_CHEB_QUICK=0 timeout 400 python bench.py --steps 3 --warmup 3 --no-c5 --no-small --no-stored --no-cpu-baseline 2>> gpurun_out/d_bench.err ) > gpurun_out/d_bench_full.json
tail -12 gpurun_out/d_pytest.log
python - <<'PY'
import json
for f in ("d_bench_quick", "d_bench_full"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["impl_config"]["mean_matvecs_per_point"], [(e["class"], round(e["share_of_step"] or 0, 3), round(e["ms_per_launch"], 4)) for e in d["roofline_path"][:5]])
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -5 gpurun_out/d_bench.err
