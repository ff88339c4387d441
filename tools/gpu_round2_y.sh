#!/bin/bash
# session Y (1 GPU): filter-knob sweep on C2 and C3 (no code change; env read at context creation)
timeout 300 python tools/probe_small_knobs.py c2 2>&1 | grep -v "^\[bh\]" | tail -12
timeout 400 python tools/probe_small_knobs.py c3 2>&1 | grep -v "^\[bh\]" | tail -12
