#!/bin/bash
# Time-boxed training on the real FB15k-237 data (shipped gcn_block.exp), then ranking of the FULL test set.
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
timeout 200 python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp \
    --dataset-npz .scratch/fb15k237_full.npz --time-budget 95 --prefetch 8 --no-periodic-eval --final-eval 0 \
    > gpurun_out/r1_fb15k237_train_full.log 2>&1
echo "train rc=$?"; tail -2 gpurun_out/r1_fb15k237_train_full.log
