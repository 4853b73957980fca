#!/bin/bash
# round-2 call 11: register prefetch distance of the GEMM producers (RGCN_GEMM_PF = 2 | 3 | 4)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_rank.py -m gpu -q -x > gpurun_out/r2c11_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2c11_gpu_tests.txt
tail -3 gpurun_out/r2c11_gpu_tests.txt
for pf in 2 3 4; do
  RGCN_GEMM_PF=$pf timeout 300 python bench.py --scale 0.1 --steps 10 --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/r2c11_syn0.1_pf$pf.json 2> gpurun_out/r2c11_syn0.1_pf$pf.err
done
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c11_full.json 2> gpurun_out/r2c11_full.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c11_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.05})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2c11_full.err
