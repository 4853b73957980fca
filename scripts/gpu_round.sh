#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list.  Everything logs into gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_fb.json 2> gpurun_out/bench_fb.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_fb.err; cat gpurun_out/bench_fb.json
timeout 600 python bench.py --workload synthetic --scale 0.02 --no-cpu-baseline --steps 5 > gpurun_out/bench_syn.json 2> gpurun_out/bench_syn.err; echo "bench syn rc=$?"
tail -3 gpurun_out/bench_syn.err; cat gpurun_out/bench_syn.json
