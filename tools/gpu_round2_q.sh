#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_small.py tests/test_thermal.py tests/test_shim.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/q_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/q_pytest.log | tail -5
for P in 8 16 32; do
( timeout 600 python bench.py --steps 3 --warmup 2 --points-per-step $P --no-c5 --no-stored --no-cpu-baseline 2> gpurun_out/q_bench_p$P.err ) > gpurun_out/q_bench_p$P.json
done
python - <<'PY'
import json
for P in (8, 16, 32):
    try:
        d = json.loads([l for l in open(f"gpurun_out/q_bench_p{P}.json") if l.startswith("{")][-1])
        print(P, round(d["value"], 3), round(d["e2e"]["value"], 3), {k: round(v["value"], 1) for k, v in (d["small_configs"] or {}).items()},
              [(e["class"], round(e["share_of_step"] or 0, 3)) for e in d["roofline_path"][:6]])
    except Exception as ex:
        print(P, "failed", ex)
PY
