#!/bin/bash
# one long FB15k-237 run WITHOUT early stopping (validation every 2000 iterations is still logged): where does the
# validation / test MRR go when training continues past the reference rule's first dip?
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
log=gpurun_out/r2_mrr_canonical_seed0_long.log
timeout 700 python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp --dataset-npz .scratch/fb15k237_full.npz \
    --time-budget ${TIME_BUDGET:-330} --prefetch 12 --final-eval 0 --seed 0 --no-save --no-early-stopping > $log 2>&1
echo "rc=$?"; grep -E "Validation" $log | tail -30; tail -1 $log
