#!/bin/bash
# call 2: tests, two-stream A/B on the graph step, conv kernel times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2m
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -4 ${O}_pytest_gpu.txt
for ts in 0 1; do
  B200_TWO_STREAM=$ts timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2_ts$ts.json 2> ${O}_bench_cfg2_ts$ts.err
  python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2_ts$ts.json').read().strip().splitlines()[-1])
    print('two_stream=$ts', d['ms_per_step'], 'ms graph;', d['config'].get('eager_ms_per_step'), 'ms eager; gemm frac', d['roofline']['frac'])
except Exception as e:
    print('bench ts=$ts failed', e); print(open('${O}_bench_cfg2_ts$ts.err').read()[-1500:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_conv.csv python tools/prof_ops.py conv 3 > /dev/null 2>&1
grep -E "dwconv" ${O}_launches_conv.csv | awk -F'","' '{print $5, $NF}' | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv -c 2 -o ${O}_dwconv python tools/prof_ops.py conv 1 > ${O}_ncu_conv.log 2>&1
