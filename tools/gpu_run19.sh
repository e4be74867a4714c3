#!/bin/bash
# call 9: evidence — full pytest, official-format bench lines (cfg2 + reference arm), one-step launch list (single stream), full ncu of the new kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2t
timeout 900 python -m pytest tests -m gpu -q --durations=8 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -3 ${O}_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_BENCH_cfg2_1gpu.json 2> ${O}_BENCH_cfg2_1gpu.err; tail -c 1200 ${O}_BENCH_cfg2_1gpu.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > ${O}_BENCH_reference_arm.json 2> ${O}_BENCH_reference_arm.err; tail -c 600 ${O}_BENCH_reference_arm.json
B200_TWO_STREAM=0 B200_GEMM_TRACE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_step_cfg2.csv python tools/step_once.py 2 > ${O}_ncu_step.log 2> ${O}_gemm_trace_all.txt
python tools/launch_summary.py ${O}_launches_step_cfg2.csv 40 > ${O}_launch_summary.txt 2>&1; head -30 ${O}_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc64|attn_bwd_tc|hc_width|dwconv" -c 8 -o ${O}_hot_ops python tools/prof_ops.py hc,conv,attn 1 > ${O}_ncu_ops.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 4 -o ${O}_gemm_epi python tools/prof_ops.py gemm 1 > ${O}_ncu_gemm.log 2>&1
ls -la gpurun_out | grep r2t
