#!/bin/bash
# GPU session U (NP GPUs): peer-memory form of the partitioned H.v: correctness and solve time vs the halo-exchange form
mkdir -p gpurun_out
NP=${NP:-2}
( timeout 600 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -8 ) > gpurun_out/u_pytest.log
grep -E "passed|failed|FAILED|rror" gpurun_out/u_pytest.log | tail -4
run() {
  ( timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/eigs_mgpu.py -m 14 -n 14 -U 4 --nev 2 --ncv 12 --check ) 2>&1 | grep -E "^\{|rror|arena|halo plan" | tail -4
}
run BH_DIST_VERBOSE=1
run BH_DIST_PEER=0
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/eigs_mgpu.py -m 12 -n 12 -U 8 --nev 20 --point --check ) 2>&1 | grep -E "^\{|rror" | tail -2
