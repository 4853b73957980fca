#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
RGCN_FUSE_DW_S5=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "rel-major" > gpurun_out/pytest_gpu_f5.log 2>&1; echo "pytest fused5 rc=$?" >> gpurun_out/pytest_gpu_f5.log
tail -5 gpurun_out/pytest_gpu_f5.log
for cfg in "4 0" "4 1" "2 0" "0 0"; do set -- $cfg
  RGCN_REL_GROUP=$1 RGCN_FUSE_DW_S5=$2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fb_g$1_f$2.json 2> gpurun_out/bench_fb.err; echo "bench rc=$?"
  tail -3 gpurun_out/bench_fb.err; python scripts/show_bench.py gpurun_out/bench_fb_g$1_f$2.json
done
