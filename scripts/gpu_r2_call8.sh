#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c8_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2c8_gpu_tests.txt
tail -8 gpurun_out/r2c8_gpu_tests.txt
timeout 300 python bench.py --scale 0.02 --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/r2c8_syn0.02.json 2> gpurun_out/r2c8_syn0.02.err
( time timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c8_full_default.json 2> gpurun_out/r2c8_full_default.err ) 2> gpurun_out/r2c8_full_default.time
timeout 600 python bench.py --workload fb15k237 --steps 50 > gpurun_out/r2c8_fb.json 2> gpurun_out/r2c8_fb.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c8_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.05}, "e2e", j.get("e2e"), "cpu", j.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/r2c8_full_default.time; tail -3 gpurun_out/r2c8_full_default.err
