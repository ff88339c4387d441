#!/bin/bash
# session W (2 GPUs): the partitioned-solve tests over all three transports
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -5
