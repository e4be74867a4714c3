#!/bin/bash
# call 3: GEGLU epilogue v2, conv tap staging, hc_width_bwd 2 blocks/SM (D<=256) + one-wave grid
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2n
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -4 ${O}_pytest_gpu.txt
timeout 300 python tools/gemm_epi_bench.py 2>&1 | tee ${O}_gemm_epi_bench.txt | head -9
B200_HC16=na timeout 300 python tools/hc_bench.py 2>&1 | tee ${O}_hc_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_ops.csv python tools/prof_ops.py conv,hc 3 > /dev/null 2>&1
grep -E "dwconv|hc_width|hc_depth" ${O}_launches_ops.csv | awk -F'","' '{print $5, $NF}' | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err
python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], 'ms graph;', d['config'].get('eager_ms_per_step'), 'ms eager; gemm frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2.err').read()[-1500:])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv -c 2 -o ${O}_dwconv python tools/prof_ops.py conv 1 > ${O}_ncu_conv.log 2>&1
