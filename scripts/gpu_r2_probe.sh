#!/bin/bash
# round-2 first GPU call: box probes, data-path microbenchmarks, the A/B of the kernels written blind in round 1,
# and an ncu capture of the HBM-regime kernels (synthetic x0.02: d=512, s=8)
mkdir -p gpurun_out
{
  echo "== box"; nproc; free -g | head -3; nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv
  python - <<'PY'
import os
print("cpu_count", os.cpu_count())
try:
    print("affinity", len(os.sched_getaffinity(0)))
except Exception as e:
    print(e)
PY
  ulimit -l
} > gpurun_out/r2_box.txt 2>&1
.scratch/bin/microbench2 > gpurun_out/r2_microbench2.txt 2>&1
bash scripts/gpu_ab_lean.sh > gpurun_out/r2_ab_lean.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_block_rel|k_gemm' -c 5 -o gpurun_out/r2_prof_syn \
  python bench.py --workload synthetic --scale 0.02 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_prof_syn.log 2>&1
ncu -i gpurun_out/r2_prof_syn.ncu-rep --page raw --csv > gpurun_out/r2_prof_syn_raw.csv 2>/dev/null
tail -5 gpurun_out/r2_microbench2.txt; tail -30 gpurun_out/r2_ab_lean.txt; cat gpurun_out/r2_box.txt
