#!/bin/bash
# session Z4 (4 GPUs): the SCALE path at N = 4 with the final solver settings (sweep + row-partitioned C5 with its check)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus 4 --no-small --no-stored --no-cpu-baseline > gpurun_out/z4_bench_n4.json 2> gpurun_out/z4_bench_n4.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/z4_bench_n4.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'checks',d.get('checks'))
    print('c5_partitioned',d.get('c5_partitioned'))
else:
    print(open('gpurun_out/z4_bench_n4.err').read()[-1500:])
PY
