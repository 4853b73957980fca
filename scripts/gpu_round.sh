#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_fb.json 2> gpurun_out/bench_fb.err; tail -3 gpurun_out/bench_fb.err; python scripts/show_bench.py gpurun_out/bench_fb.json
timeout 600 python bench.py --workload synthetic --scale 0.02 --no-cpu-baseline --steps 5 > gpurun_out/bench_syn.json 2> gpurun_out/bench_syn.err; tail -3 gpurun_out/bench_syn.err; python scripts/show_bench.py gpurun_out/bench_syn.json
timeout 600 python scripts/bench_secondary.py > gpurun_out/secondary_r1.json 2> gpurun_out/secondary.err; tail -2 gpurun_out/secondary.err
python scripts/gemm_perf.py
