#!/bin/bash
# validation of the round-2 main library: full GPU suite, bench lines (cfg2 with CPU baseline, cfg3/4/5), launch list + ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_cfg2.json 2> gpurun_out/r2g_bench_cfg2.err; tail -c 700 gpurun_out/r2g_bench_cfg2.json
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2g_bench_cfg4.json 2> gpurun_out/r2g_bench_cfg4.err; tail -c 200 gpurun_out/r2g_bench_cfg4.json
timeout 400 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu > gpurun_out/r2g_bench_cfg3.json 2> gpurun_out/r2g_bench_cfg3.err; tail -c 200 gpurun_out/r2g_bench_cfg3.json
timeout 400 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2g_bench_cfg5.json 2> gpurun_out/r2g_bench_cfg5.err; tail -c 200 gpurun_out/r2g_bench_cfg5.json
# launch list of one eager step (all kernels, device time) + DRAM bytes of the GEMM family
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/r2g_ncu_launches.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/r2g_gemm_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/r2g_ncu_traffic.log 2>&1
