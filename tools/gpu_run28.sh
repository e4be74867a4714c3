#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2f4b
timeout 900 python -m pytest tests -m gpu -q --durations=4 > ${O}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.txt
tail -25 ${O}_pytest_gpu.txt
