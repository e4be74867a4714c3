#!/bin/bash
# call: double-buffered TMA staging in the GEMM epilogue; DurationPredictor step through GraphedTrainStep (cfg4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2w
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -3 ${O}_pytest_gpu.txt
timeout 300 python tools/gemm_epi_bench.py 2>&1 | tee ${O}_gemm_epi_bench.txt
for c in 2 4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg$c.json 2> ${O}_bench_cfg$c.err
  python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg$c.json').read().strip().splitlines()[-1])
    print('cfg$c', round(d['ms_per_step'],2), 'ms;', round(d['value']), '; e2e', round(d['e2e']['value']), '; eager', d['config'].get('eager_ms_per_step'), d['config'].get('cuda_graph'), '; gemm frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
except Exception as e:
    print('bench cfg$c failed', e); print(open('${O}_bench_cfg$c.err').read()[-2500:])
PY
done
