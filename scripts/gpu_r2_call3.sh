#!/bin/bash
# round-2 GPU call 3: staged-kernel variants (TMA vs cp.async, quads per lane), the new bench.py at full C5 size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_staged.py -x -q > gpurun_out/r2c3_staged_tests.txt 2>&1
echo "staged tests rc=$?" >> gpurun_out/r2c3_staged_tests.txt
for cfg in "2 0" "3 1" "4 2" "0 3"; do
  set -- $cfg
  RGCN_BLOCK_ALGO=3 RGCN_STG_FWD=$1 RGCN_STG_BWD=$2 timeout 300 python bench.py --scale 0.02 --steps 20 --no-cpu-baseline --no-e2e \
    > gpurun_out/r2c3_syn002_f$1_b$2.json 2> gpurun_out/r2c3_syn002_f$1_b$2.err
  RGCN_BLOCK_ALGO=3 RGCN_STG_FWD=$1 RGCN_STG_BWD=$2 timeout 300 python bench.py --scale 0.1 --steps 10 --no-cpu-baseline --no-e2e \
    > gpurun_out/r2c3_syn01_f$1_b$2.json 2> gpurun_out/r2c3_syn01_f$1_b$2.err
done
# the full configuration: 10 M nodes / 100 M edges on one GPU, e2e and CPU baseline included
( time RGCN_BLOCK_ALGO=3 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c3_full.json 2> gpurun_out/r2c3_full.err ) 2> gpurun_out/r2c3_full.time
nvidia-smi --query-gpu=memory.used --format=csv >> gpurun_out/r2c3_full.time
RGCN_BLOCK_ALGO=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_block_stg' -c 2 -o gpurun_out/r2c3_prof_stg \
  python bench.py --scale 0.02 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c3_prof_stg.log 2>&1
ncu -i gpurun_out/r2c3_prof_stg.ncu-rep --page raw --csv > gpurun_out/r2c3_prof_stg_raw.csv 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c3_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.15}, j.get("e2e"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -5 gpurun_out/r2c3_staged_tests.txt; cat gpurun_out/r2c3_full.time; tail -5 gpurun_out/r2c3_full.err
