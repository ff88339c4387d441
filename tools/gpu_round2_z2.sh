#!/bin/bash
# session Z2 (1 GPU): stall rule needs >= 80 filtered steps; C5 solve time against the cut fraction; full tests; bench
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/probe_c5.py 0.05 0.08 2>&1 | grep -v "^\[bh\]" | tail -6
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/z2_bench.json 2> gpurun_out/z2_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/z2_bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'checks',d.get('checks'))
    print('small',{k:(v.get('value'),v.get('matches_reference_phase_txt')) for k,v in d.get('small_configs',{}).items()})
    print('c5',d['c5_matrix_free_hv']['ground_state_and_gap'])
else:
    print(open('gpurun_out/z2_bench.err').read()[-1500:])
PY
