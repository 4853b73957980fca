#!/bin/bash
# round-2 call 12: persistent NT GEMM with two accumulator sets in TMEM (tile t+1 MMAs overlap tile t epilogue)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_rank.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2c12_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2c12_gpu_tests.txt
tail -5 gpurun_out/r2c12_gpu_tests.txt
timeout 300 python bench.py --scale 0.02 --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/r2c12_syn0.02.json 2> gpurun_out/r2c12_syn0.02.err
RGCN_GEMM_CTAS=100000000 timeout 300 python bench.py --scale 0.02 --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/r2c12_syn0.02_onetile.json 2> gpurun_out/r2c12_syn0.02_onetile.err
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c12_full.json 2> gpurun_out/r2c12_full.err
timeout 600 python bench.py --workload fb15k237 --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/r2c12_fb.json 2> gpurun_out/r2c12_fb.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c12_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.05})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2c12_full.err
