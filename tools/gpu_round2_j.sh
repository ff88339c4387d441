#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/j_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/j_pytest.log | tail -12
BH_BATCH_VERBOSE=1 python - > gpurun_out/j_small_time.log 2>&1 <<'PY'
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
    c.close()
PY
cat gpurun_out/j_small_time.log
