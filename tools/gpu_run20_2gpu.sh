#!/bin/bash
# 2-GPU call: NCCL data-parallel correctness, N=1 vs N=2 bench on the same box, step timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2u
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q > ${O}_pytest_gpu_ddp_2rank.txt 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu_ddp_2rank.txt
tail -4 ${O}_pytest_gpu_ddp_2rank.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > ${O}_BENCH_cfg2_1gpu_same_box.json 2> ${O}_bench_n1.err; python -c "import json;d=json.loads(open('${O}_BENCH_cfg2_1gpu_same_box.json').read().strip().splitlines()[-1]);print('N=1',d['ms_per_step'],d['value'])"
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_BENCH_cfg2_2gpu.json 2> ${O}_bench_n2.err; python -c "import json;d=json.loads(open('${O}_BENCH_cfg2_2gpu.json').read().strip().splitlines()[-1]);print('N=2',d['ms_per_step'],d['value'])" || tail -20 ${O}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_timeline.py 20 2> ${O}_timeline.err | tee ${O}_ddp_timeline_2gpu.txt
tail -3 ${O}_timeline.err
