#!/bin/bash
# call 7: hc stats reuse in backward, conv fwd two tiles/block with prefetch, graph-replay roofline timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2r
timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -5 ${O}_pytest_gpu.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err
python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], 'ms graph;', d['config'].get('eager_ms_per_step'), 'ms eager; gemm frac', d['roofline']['frac'], 'gemm ms', d['roofline']['ms_per_step'], 'launches', d['gpu_launches'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2.err').read()[-2500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_ops.csv python tools/prof_ops.py conv,hc 3 > /dev/null 2>&1
grep -E "dwconv|hc_width|hc_depth" ${O}_launches_ops.csv | awk -F'","' '{print $5, $NF}' | tail -8
