#!/bin/bash
# GPU session R (1 GPU): final validation: smoke, full GPU test-suite, bench line, launch list, sanitizer on the small solver
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r_smoke.log; cat gpurun_out/r_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/r_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r_pytest.log | tail -12
( timeout 900 python bench.py --steps 4 --warmup 3 2> gpurun_out/r_bench.err ) > gpurun_out/r_bench.json
( timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r_launches.csv \
    python bench.py --steps 1 --warmup 3 --points-per-step 4 --no-cpu-baseline --no-c5 --no-small --no-stored > gpurun_out/r_ncu_bench.log 2>&1 )
( SAN_POINTS=8 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r_memcheck_small.log 2>&1 )
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_compress_tiled8 -s 5 -c 1 -f -o gpurun_out/r_restart python tools/ncu_targets.py c3 > gpurun_out/r_ncu_restart.log 2>&1 )
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r_bench.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "mv", d["impl_config"]["mean_matvecs_per_point"])
    print([(e["class"], round(e["share_of_step"] or 0, 3), round(e["ms_per_launch"], 4), round(e["frac"], 3)) for e in d["roofline_path"]])
    print("small", {k: (v.get("value"), v.get("matches_reference_phase_txt")) for k, v in d["small_configs"].items()})
    print("stored", d["stored_kernel"]["value"], "c4", d["c4_rect_4x3"]["stored"]["frac_of_hbm_peak"], "c5", d["c5_matrix_free_hv"]["ground_state_and_gap"], "checks", d["checks"], "cpu", d["cpu_baseline"]["value"])
except Exception as ex:
    print("bench failed", ex)
PY
tail -3 gpurun_out/r_bench.err; tail -4 gpurun_out/r_memcheck_small.log
