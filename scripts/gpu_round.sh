#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -x > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_train.log
tail -12 gpurun_out/pytest_train.log
timeout 600 python scripts/bench_secondary.py > gpurun_out/secondary_r1.json 2> gpurun_out/secondary.err; tail -3 gpurun_out/secondary.err
python - <<'PY'
import json; j=json.load(open('gpurun_out/secondary_r1.json'))
for k,v in j.items():
    if 'stages_ms' in v: print(k, round(v['fwd_bwd_ms'],3), v['stages_ms'])
PY
timeout 900 python bench.py --workload synthetic --scale 0.3 --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/bench_syn03.json 2> gpurun_out/bench_syn03.err; tail -3 gpurun_out/bench_syn03.err; python scripts/show_bench.py gpurun_out/bench_syn03.json
