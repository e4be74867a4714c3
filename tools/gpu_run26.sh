#!/bin/bash
# final validation of the tree: full GPU tests, smoke(), default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2final
timeout 900 python -m pytest tests -m gpu -q --durations=5 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -3 ${O}_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee ${O}_smoke.txt
timeout 600 python bench.py > ${O}_BENCH_cfg2_1gpu.json 2> ${O}_bench.err
python - <<PY
import json
d = json.loads(open('${O}_BENCH_cfg2_1gpu.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'ms;', round(d['value']), d['unit'], '; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], '; cpu', round(d['cpu_baseline']['value']), '; clocks', d['clocks'])
PY
