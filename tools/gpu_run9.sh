#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log
B200_GEMM_BREAKDOWN=1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu --no-graph > gpurun_out/r2i_bench_eager.json 2> gpurun_out/r2i_gemm_breakdown.txt; grep -A14 "GEMM breakdown" gpurun_out/r2i_gemm_breakdown.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2i_bench_cfg2.json 2> gpurun_out/r2i_bench_cfg2.err; tail -c 500 gpurun_out/r2i_bench_cfg2.json
timeout 400 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2i_bench_cfg5.json 2> gpurun_out/r2i_bench_cfg5.err; tail -c 200 gpurun_out/r2i_bench_cfg5.json
timeout 300 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu > gpurun_out/r2i_bench_cfg3.json 2> gpurun_out/r2i_bench_cfg3.err; tail -c 200 gpurun_out/r2i_bench_cfg3.json
timeout 400 ncu --set full --clock-control none -k regex:"hc_depth|dwconv|geglu_bwd|qkv_post|rowgate_bwd|hc_width" -s 60 -c 14 -o gpurun_out/r2i_small_ops python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/r2i_ncu_small.log 2>&1
