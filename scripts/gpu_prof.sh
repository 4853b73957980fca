#!/bin/bash
set -x
mkdir -p gpurun_out
./scripts/microbench > gpurun_out/microbench.txt 2>&1
cat gpurun_out/microbench.txt
# launch list (cold-cache, serialised) of the default bench
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the three hot kernels (one launch each, after warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_block_agg|k_block_dw' -s 6 -c 3 -o gpurun_out/prof_block_r1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
