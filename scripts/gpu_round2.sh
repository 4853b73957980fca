#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x > gpurun_out/pytest_parallel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parallel.log
tail -4 gpurun_out/pytest_parallel.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_fb_2gpu.json 2> gpurun_out/bench_fb_2gpu.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/bench_fb_2gpu.err | tail -3; python scripts/show_bench.py gpurun_out/bench_fb_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-200
