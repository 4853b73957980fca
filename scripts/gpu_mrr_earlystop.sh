#!/bin/bash
# FB15k-237, shipped gcn_block.exp, the reference's own stopping rule (validation filtered MRR every 2000
# iterations, burn-in 6000), then the full test set.  Needs .scratch/fb15k237_full.npz (scripts/pack_dataset.py).
# Bounded by TIME_BUDGET seconds of training (default 600) in case the criterion never fires.
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
timeout $(( ${TIME_BUDGET:-600} + 240 )) python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp \
    --dataset-npz .scratch/fb15k237_full.npz --time-budget ${TIME_BUDGET:-600} --prefetch 8 --final-eval 0 \
    > gpurun_out/fb15k237_train_earlystop.log 2>&1
echo "train rc=$?"; grep -E "Validation|Stopping|Ignoring" gpurun_out/fb15k237_train_earlystop.log | tail -20
tail -1 gpurun_out/fb15k237_train_earlystop.log
