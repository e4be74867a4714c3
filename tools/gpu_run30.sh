#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r2final2_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final2_pytest_gpu.txt
tail -3 gpurun_out/r2final2_pytest_gpu.txt
