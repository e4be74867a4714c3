#!/bin/bash
# call 5: bench after the L1 prefetch + robust roofline pass; per-kernel launch list of the step (single stream)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2p
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err
python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], 'ms graph;', d['config'].get('eager_ms_per_step'), 'ms eager; gemm frac', d['roofline']['frac'], 'gemm ms', d['roofline']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2.err').read()[-1500:])
PY
B200_TWO_STREAM=0 B200_GEMM_TRACE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_bench_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > ${O}_ncu_bench.log 2> ${O}_gemm_trace.txt
python tools/launch_summary.py ${O}_launches_bench_cfg2.csv > ${O}_launch_summary.txt 2>&1; head -32 ${O}_launch_summary.txt
timeout 200 python tools/gemm_epi_bench.py 2>&1 | tee ${O}_gemm_epi_bench.txt | sed -n 9,14p
