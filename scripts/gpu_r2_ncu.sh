#!/bin/bash
# round-2 ncu call: --set full of the team aggregation kernels and the tcgen05 GEMMs at $1 x the C5 graph (default 0.3:
# 3 M nodes / 30 M edges, H = 6 GB -- far beyond L2 like the full graph, but cheap to replay)
SC=${1:-0.3}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_block_team|k_gemm_t" --launch-skip 5 --launch-count 5 \
  -o gpurun_out/r2_ncu_x$SC -f python bench.py --scale $SC --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_x$SC.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_x$SC.log
ncu -i gpurun_out/r2_ncu_x$SC.ncu-rep --page raw --csv > gpurun_out/r2_ncu_x${SC}_raw.csv 2>/dev/null
ls -la gpurun_out/r2_ncu_x$SC*
