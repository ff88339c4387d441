#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_accel_parity.py tests/test_gpu_golden.py tests/test_gpu_sweep_parity.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/n_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/n_pytest.log | tail -5
for ring in 1 0; do
( BH_COOP_RING=$ring timeout 400 python bench.py --steps 3 --warmup 3 --no-c5 --no-small --no-stored --no-cpu-baseline 2> gpurun_out/n_bench_ring$ring.err ) > gpurun_out/n_bench_ring$ring.json
done
python - <<'PY'
import json
for f in ("n_bench_ring1", "n_bench_ring0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 3), round(d["e2e"]["value"], 3), [(e["class"], round(e["share_of_step"] or 0, 3), round(e["ms_per_launch"], 4), round(e["frac"], 3)) for e in d["roofline_path"][:5]])
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -3 gpurun_out/n_bench_ring1.err
