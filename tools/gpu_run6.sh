#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/e2-tts-pytorch_b200
timeout 200 python tools/dbg_batch_linearity.py > gpurun_out/r2f_dbg_linearity.txt 2>&1; tail -14 gpurun_out/r2f_dbg_linearity.txt
B200_GEMM_BREAKDOWN=1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu --no-graph > gpurun_out/r2f_bench_eager.json 2> gpurun_out/r2f_gemm_breakdown.txt; grep -A40 "GEMM breakdown" gpurun_out/r2f_gemm_breakdown.txt | head -45
export B200_LIB=$L/libb200e2tts_attn3.so
timeout 120 python tools/attn_bench.py cfg2 15 2>&1 | tail -2 | tee -a gpurun_out/r2f_attn_bench.txt
timeout 120 python tools/attn_bench.py cfg3 8 2>&1 | tail -2 | tee -a gpurun_out/r2f_attn_bench.txt
timeout 600 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -m gpu -q -k "attention or e2tts or sample" > gpurun_out/r2f_pytest_attn3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_attn3.log
tail -4 gpurun_out/r2f_pytest_attn3.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2f_bench_cfg2_attn3.json 2> gpurun_out/r2f_bench_cfg2_attn3.err; tail -c 300 gpurun_out/r2f_bench_cfg2_attn3.json
