#!/bin/bash
# session X (1 GPU): racecheck over the many-point small-system solver (8 points at m = n = 8)
SAN_POINTS=8 timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py 2>&1 | tail -12
