#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/e_quick.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
for m in (10, 12):
    ctx = pkg.Context(0).setup(m, m)
    for U in (4.0, 16.0):
        try:
            r = ctx.eigs(1.0, U, 1.0, nev=20, kernel=pkg.capi.HV_MATRIX_FREE, order=pkg.capi.LEX, allow_noconv=True)
            print(m, U, "rc", r["rc"], "nmatvec", r["nmatvec"], "nrestart", r["nrestart"], "sec", r["seconds"], r["evals"][:3], flush=True)
        except Exception as ex:
            print(m, U, "EXC", ex, flush=True)
    try:
        ctx.set_batch(4)
        U = np.array([1.0, 4.0, 9.0, 16.0, 25.0, 32.0])
        o3, infos = ctx.points(np.ones(6), U, np.zeros(6), kernel=pkg.capi.HV_MATRIX_FREE)
        print(m, "points ok", o3[:, 0], [i["nmatvec"] for i in infos], flush=True)
    except Exception as ex:
        print(m, "points EXC", ex, flush=True)
    ctx.close()
PY
( cd /tmp && mkdir -p optb && cd optb && OMP_NUM_THREADS=2 timeout 120 $OLDPWD/oracle/_ref/optionB_QuantumProject -m 5 -n 5 -J 1 -U 0 -u 0 -r 2 -s 1 -f J -t exact; echo "rc=$?"; ls; cat phase.txt ) > gpurun_out/e_optb.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_shim.py -m gpu -q -k "matop or phase_m10 or option_b" 2>&1 | tail -150 ) > gpurun_out/e_pytest.log
cat gpurun_out/e_quick.log; tail -30 gpurun_out/e_optb.log
