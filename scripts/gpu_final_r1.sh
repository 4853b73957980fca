#!/bin/bash
# Round-1 closing run on one B200: (1) time-boxed training on the real FB15k-237 data (packed in .scratch/)
# with the shipped gcn_block.exp, final ranking on 2000 test triples; (2) the GPU test suite; (3) a short bench.
mkdir -p gpurun_out
python - <<'PY'
import json
t = json.load(open("tests/golden/toy_golden.json"))
open("gpurun_out/gcn_block.exp", "w").write(t["settings_text"]["gcn_block.exp"])
PY
timeout 170 python -m relationprediction_b200.train --settings gpurun_out/gcn_block.exp \
    --dataset-npz .scratch/fb15k237_full.npz --time-budget 80 --prefetch 8 --no-periodic-eval --final-eval 2000 \
    > gpurun_out/r1_fb15k237_train.log 2>&1
echo "train rc=$?"; tail -4 gpurun_out/r1_fb15k237_train.log
python -m pytest tests -q -m gpu -x 2>&1 | tail -6
python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r1_final_bench.json 2> gpurun_out/r1_final_bench.err
echo "bench rc=$?"; cat gpurun_out/r1_final_bench.json
