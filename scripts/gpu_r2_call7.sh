#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_staged.py -x -q -k "default-800-2-20000" > gpurun_out/r2c7_sanitizer.txt 2>&1
grep -v "^=========     Host Frame\|^=========         in \|^=========     at " gpurun_out/r2c7_sanitizer.txt | head -60
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c7_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2c7_gpu_tests.txt
tail -12 gpurun_out/r2c7_gpu_tests.txt
for sc in 0.02 1; do
  timeout 600 python bench.py --scale $sc --steps 5 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/r2c7_syn${sc}.json 2> gpurun_out/r2c7_syn${sc}.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c7_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.15})
    except Exception as e:
        print(f, "failed", e)
PY
