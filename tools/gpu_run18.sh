#!/bin/bash
# call 8: two-CTA-per-SM attention forward (64-key tiles) A/B, shared maskbits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2s
timeout 600 python -m pytest tests -m gpu -x -q -k "attention or e2tts or sample or full_size" > ${O}_pytest_attn.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_attn.txt
tail -4 ${O}_pytest_attn.txt
for v in 0 1; do
  B200_ATTN_FWD64=$v timeout 120 python tools/attn_bench.py cfg2 15 2>&1 | tail -2 | sed "s/^/fwd64=$v /" | tee -a ${O}_attn_bench.txt
  B200_ATTN_FWD64=$v timeout 120 python tools/attn_bench.py cfg3 8 2>&1 | tail -2 | sed "s/^/fwd64=$v /" | tee -a ${O}_attn_bench.txt
done
for v in 0 1; do
  B200_ATTN_FWD64=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_bench_cfg2_a$v.json 2> ${O}_bench_cfg2_a$v.err
  python - <<PY
import json
try:
    d = json.loads(open('${O}_bench_cfg2_a$v.json').read().strip().splitlines()[-1])
    print('fwd64=$v', d['ms_per_step'], 'ms graph;', d['config'].get('eager_ms_per_step'), 'ms eager; gemm frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('${O}_bench_cfg2_a$v.err').read()[-2500:])
PY
done
