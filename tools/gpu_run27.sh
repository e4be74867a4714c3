#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2f4
timeout 600 python -m pytest tests -m gpu -q -k "fourier or cfg2_shape or golden" > ${O}_pytest_fourier.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_fourier.txt
tail -12 ${O}_pytest_fourier.txt
