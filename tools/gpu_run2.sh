#!/bin/bash
# round-2 GPU session 2: full parity suite (validated attention), then the persistent attention forward as an A/B library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_cfg2.json 2> gpurun_out/r2b_bench_cfg2.err; tail -c 600 gpurun_out/r2b_bench_cfg2.json
B200_LIB=$PWD/e2-tts-pytorch_b200/libb200e2tts_pdl.so B200_PDL=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_cfg2_pdl.json 2> gpurun_out/r2b_bench_cfg2_pdl.err; tail -c 300 gpurun_out/r2b_bench_cfg2_pdl.json
timeout 400 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_bench_cfg4.json 2> gpurun_out/r2b_bench_cfg4.err; tail -c 300 gpurun_out/r2b_bench_cfg4.json
export B200_LIB=$PWD/e2-tts-pytorch_b200/libb200e2tts_attn2.so
timeout 600 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -m gpu -q -k "attention or e2tts or sample" > gpurun_out/r2b_pytest_attn2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_attn2.log
tail -8 gpurun_out/r2b_pytest_attn2.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_cfg2_attn2.json 2> gpurun_out/r2b_bench_cfg2_attn2.err; tail -c 300 gpurun_out/r2b_bench_cfg2_attn2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ --csv --log-file gpurun_out/r2b_attn2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/r2b_ncu_attn2.log 2>&1
