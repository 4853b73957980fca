#!/bin/bash
# round-2 GPU call 2: validate the TMA-staged block kernels (block_algo = 3) and A/B them against the register path
mkdir -p gpurun_out
{
  echo "== sanitizer on the microbenchmark (small sizes)"
  timeout 300 compute-sanitizer --tool memcheck .scratch/bin/microbench2 65536 50000 BD 2>&1 | tail -25
  echo "== microbenchmark, full size"
  timeout 120 .scratch/bin/microbench2 2>&1
} > gpurun_out/r2_microbench2b.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_staged.py -x -q > gpurun_out/r2_staged_tests.txt 2>&1
echo "staged tests rc=$?" >> gpurun_out/r2_staged_tests.txt
if grep -q "passed" gpurun_out/r2_staged_tests.txt && ! grep -q "failed" gpurun_out/r2_staged_tests.txt; then
  for cfg in "1 0" "3 0" "3 1" "3 2"; do
    set -- $cfg
    RGCN_BLOCK_ALGO=$1 RGCN_STG_FWD=$2 timeout 300 python bench.py --workload synthetic --scale 0.02 --steps 20 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2_syn_algo$1_fwd$2.json 2> gpurun_out/r2_syn_algo$1_fwd$2.err
  done
  RGCN_BLOCK_ALGO=3 timeout 300 python bench.py --workload synthetic --scale 0.1 --steps 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2_syn01_algo3.json 2> gpurun_out/r2_syn01_algo3.err
  RGCN_BLOCK_ALGO=1 timeout 300 python bench.py --workload synthetic --scale 0.1 --steps 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2_syn01_algo1.json 2> gpurun_out/r2_syn01_algo1.err
  RGCN_BLOCK_ALGO=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_block_stg' -c 2 -o gpurun_out/r2_prof_stg \
    python bench.py --workload synthetic --scale 0.02 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_prof_stg.log 2>&1
  ncu -i gpurun_out/r2_prof_stg.ncu-rep --page raw --csv > gpurun_out/r2_prof_stg_raw.csv 2>/dev/null
else
  echo "== sanitizer on one staged test" >> gpurun_out/r2_staged_tests.txt
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_staged.py -x -q -k "700-5-9000-128" 2>&1 | tail -40 >> gpurun_out/r2_staged_tests.txt
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_syn*_algo*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms" % (j["value"], j["ms_per_step"]), {k: round(v, 3) for k, v in j["stages_ms"].items() if v > 0.15})
    except Exception as e:
        print(f, "failed", e)
PY
tail -15 gpurun_out/r2_staged_tests.txt; tail -30 gpurun_out/r2_microbench2b.txt
