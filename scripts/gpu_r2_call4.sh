#!/bin/bash
# round-2 GPU call 4: full GPU suite (cuBLAS-free library, staged + team kernels, fused ranker, slice norms), then
# team vs per-warp staged kernels at 0.02 / 0.1 / full scale
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c4_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2c4_gpu_tests.txt
for team in 1 0; do
  for sc in 0.02 0.1; do
    RGCN_BLOCK_ALGO=3 RGCN_STG_TEAM=$team timeout 300 python bench.py --scale $sc --steps 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2c4_syn${sc}_team$team.json 2> gpurun_out/r2c4_syn${sc}_team$team.err
  done
  RGCN_BLOCK_ALGO=3 RGCN_STG_TEAM=$team timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/r2c4_full_team$team.json 2> gpurun_out/r2c4_full_team$team.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c4_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if v > 0.15})
    except Exception as e:
        print(f, "failed", e)
PY
tail -25 gpurun_out/r2c4_gpu_tests.txt
