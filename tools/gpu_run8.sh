#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_HC16=0 timeout 200 python tools/hc_bench.py 2>&1 | tail -3 | tee gpurun_out/r2h_hc_bench.txt
B200_HC16=1 timeout 200 python tools/hc_bench.py 2>&1 | tail -3 | tee -a gpurun_out/r2h_hc_bench.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2h_bench_cfg2.json 2> gpurun_out/r2h_bench_cfg2.err; tail -c 500 gpurun_out/r2h_bench_cfg2.json
timeout 400 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2h_bench_cfg5.json 2> gpurun_out/r2h_bench_cfg5.err; tail -c 300 gpurun_out/r2h_bench_cfg5.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hc_width_bwd16 -s 1 -c 1 -o gpurun_out/r2h_hc_bwd16 python tools/prof_ops.py hc 2 > gpurun_out/r2h_ncu_hc.log 2>&1
