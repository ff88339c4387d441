#!/bin/bash
# GPU session L (NP GPUs): partitioned H.v grid variants, then the bench line at N = NP (SCALE path incl. c5_partitioned)
mkdir -p gpurun_out
NP=${NP:-2}
run() {
  ( timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/hv_mgpu.py 14 20 ) 2>&1 | grep -E "^\{|rror" | tail -2
}
if [ "${VARIANTS:-1}" = "1" ]; then
run BH_HALO_GRID=0
run BH_HALO_GRID=7
run BH_HALO_GRID=6
run BH_HALO_GRID=4
run BH_HALO_GRID=7 BH_HALO_ABLATE=5
run BH_DIST_ALLGATHER=1
fi
( BH_DIST_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $NP --steps 3 --warmup 3 2> gpurun_out/l_bench_n$NP.err ) > gpurun_out/l_bench_n$NP.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/l_bench_n$NP.json"))
    print("N", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "c5", d.get("c5_partitioned"))
except Exception as ex:
    print("bench failed", ex)
PY
tail -5 gpurun_out/l_bench_n$NP.err
