#!/bin/bash
# round-2 GPU session 1: parity at BASELINE shapes, bench lines for cfg2/3/4/5, PDL A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err; tail -c 1500 gpurun_out/r2a_bench_cfg2.json
B200_LIB=$PWD/e2-tts-pytorch_b200/libb200e2tts_pdl.so B200_PDL=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg2_pdl.json 2> gpurun_out/r2a_bench_cfg2_pdl.err; tail -c 600 gpurun_out/r2a_bench_cfg2_pdl.json
timeout 400 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err; tail -c 400 gpurun_out/r2a_bench_cfg4.json
timeout 600 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; tail -c 400 gpurun_out/r2a_bench_cfg3.json
timeout 600 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2a_bench_cfg5.json 2> gpurun_out/r2a_bench_cfg5.err; tail -c 400 gpurun_out/r2a_bench_cfg5.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; tail -c 400 gpurun_out/r2a_bench_ref.json
