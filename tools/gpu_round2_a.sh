#!/bin/bash
# GPU session A: tests, bench, launch list, sanitizers.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/a_pytest.log
( timeout 600 python bench.py --steps 4 --warmup 3 2> gpurun_out/a_bench.err ) > gpurun_out/a_bench.json
( timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 1 --warmup 3 --points-per-step 4 --no-cpu-baseline --no-c5 --no-small --no-stored > gpurun_out/a_ncu_bench.log 2>&1 )
( SAN_POINTS=4 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/a_memcheck.log 2>&1 )
( SAN_POINTS=2 timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/a_racecheck.log 2>&1 )
tail -5 gpurun_out/a_pytest.log
head -c 600 gpurun_out/a_bench.json
tail -3 gpurun_out/a_memcheck.log gpurun_out/a_racecheck.log
