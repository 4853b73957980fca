#!/bin/bash
# where does a training iteration go?  FB15k-237 / gcn_block.exp: phase-synchronised breakdown and free-running rate,
# with the live sampler (12 / 24 threads) and with a repeated sample (no host sampler in the loop)
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
for cfg in "--prefetch 12" "--prefetch 24" "--prefetch 12 --repeat-sample"; do
  timeout 300 python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp --dataset-npz .scratch/fb15k237_full.npz \
      --seed 0 --no-save --no-periodic-eval --profile-iterations 300 $cfg 2>&1 | tail -1
done | tee gpurun_out/r2_train_profile.txt
N=${1:-0}
if [ "$N" -gt 1 ]; then bash scripts/gpu_r2_n8b.sh $N; fi
