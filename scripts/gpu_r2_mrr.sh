#!/bin/bash
# round-2 MRR runs: FB15k-237, shipped gcn_block.exp, the reference's early-stopping rule (validation filtered MRR every
# 2000 iterations, burn-in 6000, previous-check comparison), full test set at the end.  One run per (norm mode, seed).
# Usage: gpu_r2_mrr.sh "<mode:seed> ..."   e.g.  "canonical:0 canonical:1 tf_unsorted_compat:0"
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
for spec in ${1:-canonical:0}; do
  mode=${spec%%:*}; seed=${spec##*:}
  extra=""; [ "$mode" != "canonical" ] && extra="--set Encoder.NormalizationMode=$mode"
  log=gpurun_out/r2_mrr_${mode}_seed${seed}.log
  timeout $(( ${TIME_BUDGET:-420} + 200 )) python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp \
      --dataset-npz .scratch/fb15k237_full.npz --time-budget ${TIME_BUDGET:-420} --prefetch 12 --final-eval 0 --seed $seed \
      --no-save $extra > $log 2>&1
  echo "== $spec rc=$?"; grep -E "Validation|Stopping|Ignoring" $log | tail -12; tail -1 $log
done
