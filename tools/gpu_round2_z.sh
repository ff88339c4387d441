#!/bin/bash
# session Z (1 GPU): cut fraction 0.08 -> 0.05 + early stall exit of a misplaced quick-mode cut: probe, full tests, bench
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/probe_small_knobs.py c2 0.04,0.06,0.08 2>&1 | grep -v "^\[bh\]" | tail -5
timeout 200 python tools/probe_small_knobs.py c3 0.04,0.06 2>&1 | grep -v "^\[bh\]" | tail -4
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/z_bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'checks',d.get('checks'))
    print('small',{k:(v.get('points_per_s'),v.get('matches_reference_phase_txt')) for k,v in d.get('small_configs',{}).items()})
else:
    print(open('gpurun_out/z_bench.err').read()[-1500:])
PY
