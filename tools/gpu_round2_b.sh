#!/bin/bash
# GPU session B (2 GPUs): the row-partitioned solve -- correctness vs one GPU, halo plan, timing vs the all-gather baseline.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15 ) > gpurun_out/b_pytest_dist.log
run() {  # name, env..., -- args
  name=$1; shift
  ( timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29541 \
      tools/eigs_mgpu.py -m 14 -n 14 -U 4 --nev 2 --ncv 12 --check ) > gpurun_out/b_${name}.log 2>&1
  grep -E "^\{|halo plan|Error|error" gpurun_out/b_${name}.log | tail -6
}
run halo BH_DIST_VERBOSE=1
run allgather BH_DIST_ALLGATHER=1
run halo20 BH_DIST_VERBOSE=0 NEV20=1
tail -5 gpurun_out/b_pytest_dist.log
