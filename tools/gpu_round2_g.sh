#!/bin/bash
# GPU session G (NP GPUs): partitioned H.v with the stored remote-hop matrix: correctness, part timings, solve time
mkdir -p gpurun_out
NP=${NP:-2}
( timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_boundary.py -q 2>&1 | tail -30 ) > gpurun_out/g_pytest.log
run() {
  ( timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/hv_mgpu.py 14 20 ) 2>&1 | grep -E "^\{|rror|halo plan" | tail -4
}
run BH_DIST_VERBOSE=1
run BH_DIST_ALLGATHER=1
run BH_HALO_ABLATE=6
run BH_HALO_ABLATE=5
run BH_HALO_ABLATE=3
run BH_HALO_ABLATE=1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/eigs_mgpu.py -m 14 -n 14 -U 4 --nev 2 --ncv 12 --check ) 2>&1 | grep -E "^\{|rror" | tail -3
( BH_DIST_ALLGATHER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/eigs_mgpu.py -m 14 -n 14 -U 4 --nev 2 --ncv 12 --check ) 2>&1 | grep -E "^\{|rror" | tail -3
grep -E "passed|failed|FAILED|Error" gpurun_out/g_pytest.log | tail
