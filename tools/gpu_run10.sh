#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/e2-tts-pytorch_b200
for v in libb200e2tts.so libb200e2tts_g2.so; do
  B200_LIB=$L/$v timeout 120 python tools/attn_bench.py cfg2 15 2>&1 | tail -2 | tee -a gpurun_out/r2j_attn_bench.txt
  B200_LIB=$L/$v timeout 120 python tools/attn_bench.py cfg3 8 2>&1 | tail -2 | tee -a gpurun_out/r2j_attn_bench.txt
done
export B200_LIB=$L/libb200e2tts_g2.so
timeout 600 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -m gpu -q -k "attention or e2tts or sample" > gpurun_out/r2j_pytest_g2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_g2.log
tail -4 gpurun_out/r2j_pytest_g2.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2j_bench_cfg2_g2.json 2> gpurun_out/r2j_bench_cfg2_g2.err; tail -c 300 gpurun_out/r2j_bench_cfg2_g2.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 1 -c 1 -o gpurun_out/r2j_attn_g2_fwd python tools/prof_ops.py attn 2 > gpurun_out/r2j_ncu.log 2>&1
