#!/bin/bash
# round-2 GPU call 5: full GPU suite, gather rate vs table size, SM-count sensitivity of the staged kernels at full size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c5_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2c5_gpu_tests.txt
timeout 300 .scratch/bin/microbench3 > gpurun_out/r2c5_microbench3.txt 2>&1
for sms in 148 110 80; do
  RGCN_BLOCK_ALGO=3 RGCN_STG_SMS=$sms timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2c5_full_sms$sms.json 2> gpurun_out/r2c5_full_sms$sms.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c5_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.15})
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/r2c5_microbench3.txt; tail -15 gpurun_out/r2c5_gpu_tests.txt
