#!/bin/bash
# evidence call: GEMM DRAM traffic of one step (cfg2, cfg3) for roofline.traffic, one-step launch list of cfg3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2z
B200_TWO_STREAM=0 timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file ${O}_gemm_traffic_cfg2.csv python tools/step_once.py 2 > /dev/null 2>&1
n2=$(grep -c gemm_tcgen05 ${O}_gemm_traffic_cfg2.csv); n2=$((n2 / 3))
python tools/gemm_traffic.py ${O}_gemm_traffic_cfg2.csv $n2 cfg2
B200_TWO_STREAM=0 timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_step_cfg3_all.csv python tools/step_once.py 3 > /dev/null 2>&1
grep -E '^"ID"|gpu__time_duration' ${O}_launches_step_cfg3_all.csv > ${O}_launches_step_cfg3.csv
python tools/launch_summary.py ${O}_launches_step_cfg3.csv 24 > ${O}_launch_summary_cfg3.txt 2>&1; head -16 ${O}_launch_summary_cfg3.txt
grep -E '^"ID"|gemm_tcgen05' ${O}_launches_step_cfg3_all.csv > ${O}_gemm_traffic_cfg3.csv
n3=$(grep -c gemm_tcgen05 ${O}_gemm_traffic_cfg3.csv); n3=$((n3 / 3))
python tools/gemm_traffic.py ${O}_gemm_traffic_cfg3.csv $n3 cfg3
cp profiles/r2_gemm_traffic.json ${O}_gemm_traffic.json
