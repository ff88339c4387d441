#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -120 ) > gpurun_out/f_pytest.log
( timeout 400 python bench.py --steps 3 --warmup 3 --no-c5 --no-small --no-cpu-baseline 2> gpurun_out/f_bench.err ) > gpurun_out/f_bench_quick.json
grep -E "passed|failed|FAILED" gpurun_out/f_pytest.log | tail -15
python - <<'PY'
import json
for f in ("f_bench_quick",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["impl_config"]["mean_matvecs_per_point"], [(e["class"], round(e["share_of_step"] or 0, 3), round(e["ms_per_launch"], 4)) for e in d["roofline_path"][:6]], d["stored_kernel"])
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -5 gpurun_out/f_bench.err
