#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2f4c
timeout 400 python -m pytest tests -m gpu -q -k "concat_cond or interpolated or fourier or golden or velocity" > ${O}_pytest_variants.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_variants.txt
tail -6 ${O}_pytest_variants.txt
