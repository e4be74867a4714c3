#!/bin/bash
# call 4: TMA-store GEMM epilogue
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2o
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -15 ${O}_pytest_gpu.txt
timeout 300 python tools/gemm_epi_bench.py 2>&1 | tee ${O}_gemm_epi_bench.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err
python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], 'ms graph;', d['config'].get('eager_ms_per_step'), 'ms eager; gemm frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2.err').read()[-1500:])
PY
