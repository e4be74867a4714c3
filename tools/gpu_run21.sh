#!/bin/bash
# call: the other BASELINE configs through bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2v
for c in 3 4 5; do
  timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu > ${O}_BENCH_cfg${c}_1gpu.json 2> ${O}_bench_cfg$c.err
  python - <<PY
import json
try:
    d = json.loads(open('${O}_BENCH_cfg${c}_1gpu.json').read().strip().splitlines()[-1])
    print('cfg$c', round(d['ms_per_step'],2), 'ms;', round(d['value']), d['unit'], '; e2e', round(d['e2e']['value']), '; eager', d['config'].get('eager_ms_per_step'), '; gemm frac', round(d['roofline']['frac'],3), 'step tensor frac', round(d['roofline']['step_tensor_frac'],3), 'launches', d['gpu_launches'])
except Exception as e:
    print('bench cfg$c failed', e); print(open('${O}_bench_cfg$c.err').read()[-2500:])
PY
done
