#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
tail -30 gpurun_out/pytest_gemm.log
timeout 300 python scripts/gemm_perf.py > gpurun_out/gemm_perf.txt 2>&1; cat gpurun_out/gemm_perf.txt
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_gemm.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
