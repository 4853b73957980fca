#!/bin/bash
# WN18 (BASELINE configs[2]): shipped gcn_basis.exp with B=2, d=200 (the edits BASELINE.md names), time-boxed
# training, full test set.  Needs .scratch/wn18_full.npz (scripts/pack_dataset.py /root/reference/data/wn18 .scratch/wn18_full.npz).
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_basis.exp", "w").write(t["settings_text"]["gcn_basis.exp"])
PY
timeout $(( ${TIME_BUDGET:-120} + 240 )) python -m relationprediction_b200.train --settings gpurun_out/gcn_basis.exp \
    --set Encoder.NumberOfBasisFunctions=2 --set Encoder.InternalEncoderDimension=200 --set Shared.CodeDimension=200 \
    --dataset-npz .scratch/wn18_full.npz --time-budget ${TIME_BUDGET:-120} --prefetch 8 --no-periodic-eval --final-eval 0 \
    > gpurun_out/wn18_train.log 2>&1
echo "train rc=$?"; tail -2 gpurun_out/wn18_train.log
