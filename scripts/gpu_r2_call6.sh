#!/bin/bash
# round-2 GPU call 6: dynamic work distribution in the staged kernels (full-size drift fix)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_staged.py tests/test_gpu_parity.py -q > gpurun_out/r2c6_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2c6_gpu_tests.txt
for sc in 0.02 0.1; do
  timeout 300 python bench.py --scale $sc --steps 10 --no-cpu-baseline --no-e2e \
    > gpurun_out/r2c6_syn${sc}.json 2> gpurun_out/r2c6_syn${sc}.err
done
for sms in 148 110; do
  RGCN_STG_SMS=$sms timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2c6_full_sms$sms.json 2> gpurun_out/r2c6_full_sms$sms.err
done
RGCN_STG_TEAM=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2c6_full_noteam.json 2> gpurun_out/r2c6_full_noteam.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c6_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.15})
    except Exception as e:
        print(f, "failed", e)
PY
tail -6 gpurun_out/r2c6_gpu_tests.txt; tail -3 gpurun_out/r2c6_full_sms148.err
