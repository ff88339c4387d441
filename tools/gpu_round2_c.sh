#!/bin/bash
# GPU session C (NP GPUs): timing of the partitioned H.v and of its parts
mkdir -p gpurun_out
NP=${NP:-2}
run() {
  ( timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
      tools/hv_mgpu.py 14 20 ) 2>&1 | grep -E "^\{|rror" | tail -3
}
run BH_X=full
run BH_DIST_ALLGATHER=1
run BH_HALO_ABLATE=6
run BH_HALO_ABLATE=5
run BH_HALO_ABLATE=3
run BH_HALO_ABLATE=1
run BH_X=full NCCL_MAX_P2P_NCHANNELS=32 NCCL_MIN_P2P_NCHANNELS=32
run BH_HALO_ABLATE=6 NCCL_MAX_P2P_NCHANNELS=32 NCCL_MIN_P2P_NCHANNELS=32
run BH_HALO_ABLATE=6 NCCL_P2P_NET_CHUNKSIZE=524288 NCCL_NCHANNELS_PER_NET_PEER=8
