#!/bin/bash
# GPU session T: ncu --set full of the 4-vector lockstep H.v kernel (second-largest share of a C3 step)
mkdir -p gpurun_out
( timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_hv_chain_batch" -s 400 -c 1 -f -o gpurun_out/t_hvbatch python tools/ncu_targets.py c3 > gpurun_out/t_ncu_hvbatch.log 2>&1 )
tail -3 gpurun_out/t_ncu_hvbatch.log; ls -la gpurun_out/t_hvbatch.ncu-rep
