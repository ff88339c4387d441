#!/bin/bash
# GPU session V (1 GPU): final validation of the committed tree: smoke, full GPU test-suite, bench line
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/v_smoke.log; cat gpurun_out/v_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/v_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/v_pytest.log | tail -12
( timeout 900 python bench.py --steps 4 --warmup 3 2> gpurun_out/v_bench.err ) > gpurun_out/v_bench.json
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/v_bench.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "checks", d["checks"])
    print("small", {k: (v.get("value"), v.get("matches_reference_phase_txt")) for k, v in d["small_configs"].items()}, "c5", d["c5_matrix_free_hv"]["ground_state_and_gap"]["solve_s"])
except Exception as ex:
    print("bench failed", ex)
PY
tail -3 gpurun_out/v_bench.err
