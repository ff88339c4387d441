#!/bin/bash
# GPU session M (NP GPUs): pipelined halo exchange: correctness, piece / grid variants, solve
mkdir -p gpurun_out
NP=${NP:-2}
( timeout 600 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -8 ) > gpurun_out/m_pytest.log
grep -E "passed|failed|FAILED" gpurun_out/m_pytest.log
run() {
  ( timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/hv_mgpu.py 14 20 ) 2>&1 | grep -E "^\{|rror|halo plan" | tail -3
}
run BH_DIST_VERBOSE=1
run BH_HALO_PIECES=8
run BH_HALO_PIECES=2
run BH_HALO_PIECES=1
run BH_HALO_GRID=8
run BH_HALO_GRID=6
run BH_HALO_ABLATE=1
run BH_HALO_ABLATE=2
run BH_DIST_ALLGATHER=1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/eigs_mgpu.py -m 14 -n 14 -U 4 --nev 2 --ncv 12 --check ) 2>&1 | grep -E "^\{|rror" | tail -3
