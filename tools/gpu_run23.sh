#!/bin/bash
# call: software-pipelined TMEM loads in the plain GEMM epilogue (200 registers); validation of the cleaned-up attention file
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2x
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -3 ${O}_pytest_gpu.txt
timeout 300 python tools/gemm_epi_bench.py 2>&1 | tee ${O}_gemm_epi_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err
python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2.json').read().strip().splitlines()[-1])
    print('cfg2', round(d['ms_per_step'],2), 'ms;', round(d['value']), '; e2e', round(d['e2e']['value']), '; eager', d['config'].get('eager_ms_per_step'), '; gemm frac', round(d['roofline']['frac'],3), 'gemm ms', round(d['roofline']['ms_per_step'],2), 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2.err').read()[-2500:])
PY
