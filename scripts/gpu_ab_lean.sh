#!/bin/bash
# A/B of the opt-in variants that were written without GPU access (one B200, ~2 min):
#   1. parity of the lean kernels (the xfail-guarded test reports XPASS / XFAIL per shape)
#   2. default vs RGCN_LEAN=1 on the FB15k-237 shape (s=5 group kernel) and on the synthetic shape (s=8)
mkdir -p gpurun_out
RGCN_RUN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k lean -rxX 2>&1 | tail -15
for lean in 0 1; do
  RGCN_LEAN=$lean python bench.py --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/ab_fb_lean$lean.json 2>/dev/null
  RGCN_LEAN=$lean python bench.py --workload synthetic --scale 0.02 --steps 30 --no-cpu-baseline --no-e2e \
      > gpurun_out/ab_syn_lean$lean.json 2>/dev/null
done
python - <<'PY'
import json
for name in ("fb", "syn"):
    for lean in (0, 1):
        try:
            j = json.loads(open("gpurun_out/ab_%s_lean%d.json" % (name, lean)).read().strip().splitlines()[-1])
            print(name, "lean=%d" % lean, "%.1f M-edges/s" % j["value"], "%.4f ms" % j["ms_per_step"], j.get("stages_ms"))
        except Exception as exc:
            print(name, lean, "failed:", exc)
PY
# 3. the component-major path (block_algo = 2): parity, then the FB15k-237-shape bench
RGCN_RUN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k component_major -rxX 2>&1 | tail -8
RGCN_BLOCK_ALGO=2 python bench.py --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/ab_fb_cm.json 2>/dev/null
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/ab_fb_cm.json").read().strip().splitlines()[-1])
    print("fb component-major", "%.1f M-edges/s" % j["value"], "%.4f ms" % j["ms_per_step"], j.get("stages_ms"))
except Exception as exc:
    print("component-major bench failed:", exc)
PY
# 4. stream-ordered graph destroy in the training loop (ms per iteration, 20 s each; needs .scratch/fb15k237_full.npz)
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
for af in 0 1; do
  RGCN_ASYNC_FREE=$af timeout 120 python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp \
      --dataset-npz .scratch/fb15k237_full.npz --time-budget 20 --prefetch 16 --no-periodic-eval --final-eval 500 \
      2>&1 | tail -1 | python -c "import sys, json; j = json.loads(sys.stdin.read()); print('async_free=$af', j['ms_per_iteration'], 'ms/iteration', j['filtered'])"
done
