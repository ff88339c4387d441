#!/bin/bash
# GPU session P (1 GPU): clean-build validation: smoke, full GPU test-suite, full bench line
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/p_smoke.log; cat gpurun_out/p_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/p_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/p_pytest.log | tail -12
( timeout 900 python bench.py --steps 4 --warmup 3 2> gpurun_out/p_bench.err ) > gpurun_out/p_bench.json
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/p_bench.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "mv", d["impl_config"]["mean_matvecs_per_point"])
    print("roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "share_of_step", "ms_per_launch", "traffic")})
    print("small", {k: (v.get("value"), v.get("matches_reference_phase_txt")) for k, v in d["small_configs"].items()})
    print("stored", d["stored_kernel"]); print("c4", d.get("c4_rect_4x3")); print("c5", d.get("c5_matrix_free_hv")); print("checks", d["checks"])
    print("seam", d["hv"]["matop_seam_host_vectors"]); print("cpu", d["cpu_baseline"]["value"])
except Exception as ex:
    print("bench failed", ex)
PY
tail -3 gpurun_out/p_bench.err
