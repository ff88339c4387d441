#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_small.py -q -x 2>&1 | tail -60 ) > gpurun_out/h_pytest_small.log
grep -E "passed|failed|FAILED|Error" gpurun_out/h_pytest_small.log | tail
python - > gpurun_out/h_small_time.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
p1 = 1.0 + np.arange(11.0); p2 = np.arange(11.0)
sU, smu = [a.reshape(-1) for a in np.meshgrid(p1, p2, indexing="ij")]
for m in (8, 10):
    c = pkg.Context(0).setup(m, m)
    c.set_batch(4)
    c.points(np.ones(8), sU[:8], smu[:8], kernel=pkg.capi.HV_MATRIX_FREE)
    t0 = time.perf_counter()
    o3, infos = c.points(np.ones(121), sU, smu, kernel=pkg.capi.HV_MATRIX_FREE)
    dt = time.perf_counter() - t0
    print(m, "121 points", dt, "s", 121 / dt, "points/s", "mean nmatvec", np.mean([i["nmatvec"] for i in infos]), flush=True)
    c.profile_enable(True); c.profile_read()
    c.points(np.ones(121), sU, smu, kernel=pkg.capi.HV_MATRIX_FREE)
    pr = c.profile_read(); c.profile_enable(False)
    print("   profile", {k: v for k, v in pr.items() if v["launches"]}, flush=True)
    c.close()
PY
cat gpurun_out/h_small_time.log
( cd /tmp && rm -rf clit && mkdir clit && cd clit && ( time /root/repo/bose-hubbard-phase-transition_b200/QuantumProject -m 8 -n 8 -J 1 -U 0 -u 0 -r 10 -s 1 -f J -t exact --no-plot > /dev/null ) 2>&1 | grep real; ( time /root/repo/bose-hubbard-phase-transition_b200/QuantumProject -m 10 -n 10 -J 1 -U 0 -u 0 -r 10 -s 1 -f J -t exact --no-plot > /dev/null ) 2>&1 | grep real )
