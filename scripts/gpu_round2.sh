#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x > gpurun_out/pytest_parallel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parallel.log
tail -8 gpurun_out/pytest_parallel.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/bench_parallel_breakdown.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_fb_2gpu.json 2> gpurun_out/bench_fb_2gpu.err; echo "rc=$?"; tail -3 gpurun_out/bench_fb_2gpu.err; cut -c1-200 gpurun_out/bench_fb_2gpu.json
