#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x > gpurun_out/pytest_parallel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parallel.log
tail -15 gpurun_out/pytest_parallel.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fb_1gpu.json 2> gpurun_out/bench_fb_1gpu.err; tail -3 gpurun_out/bench_fb_1gpu.err; python scripts/show_bench.py gpurun_out/bench_fb_1gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_fb_2gpu.json 2> gpurun_out/bench_fb_2gpu.err; echo "rc=$?"; tail -5 gpurun_out/bench_fb_2gpu.err; cat gpurun_out/bench_fb_2gpu.json | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --workload synthetic --scale 0.02 > gpurun_out/bench_syn_2gpu.json 2> gpurun_out/bench_syn_2gpu.err; echo "rc=$?"; tail -5 gpurun_out/bench_syn_2gpu.err; cat gpurun_out/bench_syn_2gpu.json | cut -c1-600
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
