#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/e2-tts-pytorch_b200
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -6 gpurun_out/r2c_pytest.log
for v in libb200e2tts.so libb200e2tts_attn2.so libb200e2tts_v1spin.so libb200e2tts_v2spin.so; do
  B200_LIB=$L/$v timeout 120 python tools/attn_bench.py cfg2 15 2>&1 | tail -2 | tee -a gpurun_out/r2c_attn_bench.txt
done
B200_LIB=$L/libb200e2tts.so timeout 120 python tools/attn_bench.py cfg3 8 2>&1 | tail -2 | tee -a gpurun_out/r2c_attn_bench.txt
B200_LIB=$L/libb200e2tts_attn2.so timeout 120 python tools/attn_bench.py cfg3 8 2>&1 | tail -2 | tee -a gpurun_out/r2c_attn_bench.txt
B200_LIB=$L/libb200e2tts_attn2.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 1 -c 1 -o gpurun_out/r2c_attn2_fwd python tools/prof_ops.py attn 2 > gpurun_out/r2c_ncu_attn2.log 2>&1
B200_LIB=$L/libb200e2tts.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 1 -c 1 -o gpurun_out/r2c_attn1_fwd python tools/prof_ops.py attn 2 > gpurun_out/r2c_ncu_attn1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
