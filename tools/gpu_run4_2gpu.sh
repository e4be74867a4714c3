#!/bin/bash
# 2-GPU session: NCCL data-parallel correctness + N=2 bench (graph replay + one flat all-reduce) + N=1 on the same box for the ratio
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q > gpurun_out/r2d_pytest_ddp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_ddp.log
tail -5 gpurun_out/r2d_pytest_ddp.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 300 gpurun_out/r2d_bench_n1.json
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; tail -c 600 gpurun_out/r2d_bench_n2.json
tail -5 gpurun_out/r2d_bench_n2.err
