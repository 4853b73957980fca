#!/bin/bash
# training iteration with the one-call native sample (rgcn_sampler_draw_batch) vs the numpy pipeline, then one
# early-stopping run with it (does the MRR stay where the numpy-sampled runs were?)
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
for cfg in "--prefetch 12" "--prefetch 6" "--prefetch 12 --numpy-sampling" "--prefetch 12 --repeat-sample"; do
  echo "== $cfg"
  timeout 300 python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp --dataset-npz .scratch/fb15k237_full.npz \
      --seed 0 --no-save --no-periodic-eval --profile-iterations 300 $cfg 2>&1 | tail -1
done | tee gpurun_out/r2_train_profile_native.txt
log=gpurun_out/r2_mrr_canonical_seed0_native_sampler.log
timeout 500 python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp --dataset-npz .scratch/fb15k237_full.npz \
    --time-budget 300 --prefetch 12 --final-eval 0 --seed 0 --no-save > $log 2>&1
echo "rc=$?"; grep -E "Validation|Stopping|Ignoring" $log | tail -12; tail -1 $log
