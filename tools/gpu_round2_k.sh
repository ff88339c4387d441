#!/bin/bash
# GPU session K (1 GPU): full test-suite, full bench line, launch list of the bench command, ncu --set full of the top kernels
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/k_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/k_pytest.log | tail -12
( timeout 600 python bench.py --steps 4 --warmup 3 2> gpurun_out/k_bench.err ) > gpurun_out/k_bench.json
( timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/k_launches.csv \
    python bench.py --steps 1 --warmup 3 --points-per-step 4 --no-cpu-baseline --no-c5 --no-small --no-stored > gpurun_out/k_ncu_bench.log 2>&1 )
cap() {  # name kernel-regex skip mode
  ( timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/k_$1 python tools/ncu_targets.py $4 > gpurun_out/k_ncu_$1.log 2>&1 )
  ( ncu -i gpurun_out/k_$1.ncu-rep --page raw --csv > gpurun_out/k_$1.raw.csv 2>/dev/null )
}
cap step k_step_coop 300 c3
cap hvbatch4 "k_hv_chain_batch<12, 4, 1" 300 c3
cap restart k_compress_tiled8 5 c3
cap small10 k_small_cycle 4 c2
cap hv14 k_hv_free_chain 2 c5
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/k_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "mv", d["impl_config"]["mean_matvecs_per_point"])
    print("roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "share_of_step", "ms_per_launch")})
    print("small", d["small_configs"]); print("stored", d["stored_kernel"]); print("c5", d.get("c5_matrix_free_hv")); print("checks", d["checks"])
    print("cpu", d["cpu_baseline"]["value"])
except Exception as ex:
    print("bench failed", ex)
PY
tail -3 gpurun_out/k_bench.err; ls -la gpurun_out/k_*.ncu-rep
