#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph_prep.py tests/test_gpu_model.py -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/e2e_probe.py 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/bench_fb.json 2> gpurun_out/bench_fb.err; tail -3 gpurun_out/bench_fb.err; python scripts/show_bench.py gpurun_out/bench_fb.json
python - <<'PY'
import json; j=json.loads(open('gpurun_out/bench_fb.json').read().strip().splitlines()[-1]); print(j['e2e']); print(j['roofline']); print(j['clocks'], j['wall_s_timed_region'])
PY
