#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2y
for pr in 0 -1; do
  B200_GRAPH_PRIORITY=$pr timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2_p$pr.json 2> ${O}_bench_cfg2_p$pr.err
  python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2_p$pr.json').read().strip().splitlines()[-1])
    print('prio $pr: cfg2', round(d['ms_per_step'],3), 'ms;', round(d['value']), '; e2e', round(d['e2e']['value']), d['config'].get('cuda_graph'))
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2_p$pr.err').read()[-2500:])
PY
done
