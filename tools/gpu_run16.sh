#!/bin/bash
# call 6: fused hyper-connection depth -> width kernels, no materialised zero grads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2q
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -6 ${O}_pytest_gpu.txt
for f in 0 1; do
  B200_FUSE_HC=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2_fuse$f.json 2> ${O}_bench_cfg2_fuse$f.err
  python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2_fuse$f.json').read().strip().splitlines()[-1])
    print('fuse_hc=$f', d['ms_per_step'], 'ms graph;', d['config'].get('eager_ms_per_step'), 'ms eager; gemm frac', d['roofline']['frac'], 'gemm ms', d['roofline']['ms_per_step'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2_fuse$f.err').read()[-1500:])
PY
done
