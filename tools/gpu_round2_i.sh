#!/bin/bash
mkdir -p gpurun_out
( python tools/probe_multiplicity.py; BH_ACCEL_MIN_D=100 python tools/probe_multiplicity.py; BH_COOP=0 python tools/probe_multiplicity.py ) > gpurun_out/i_mult.log 2>&1
cat gpurun_out/i_mult.log
