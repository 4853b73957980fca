#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both aggregation algorithms).  Logs into gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
for algo in 0 1; do
  RGCN_BLOCK_ALGO=$algo timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fb_algo$algo.json 2> gpurun_out/bench_fb_algo$algo.err; echo "bench rc=$?"
  tail -3 gpurun_out/bench_fb_algo$algo.err; python scripts/show_bench.py gpurun_out/bench_fb_algo$algo.json
  RGCN_BLOCK_ALGO=$algo timeout 600 python bench.py --workload synthetic --scale 0.02 --no-cpu-baseline --steps 5 > gpurun_out/bench_syn_algo$algo.json 2> gpurun_out/bench_syn_algo$algo.err; echo "bench syn rc=$?"
  tail -3 gpurun_out/bench_syn_algo$algo.err; python scripts/show_bench.py gpurun_out/bench_syn_algo$algo.json
done
