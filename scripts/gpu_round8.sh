#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x > gpurun_out/pytest_parallel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parallel.log
tail -6 gpurun_out/pytest_parallel.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fb_${n}gpu.json 2> gpurun_out/bench_fb_${n}gpu.err; echo "rc=$?"
  grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/bench_fb_${n}gpu.err | tail -4
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_fb_${n}gpu.json').read().strip().splitlines()[-1]); print('N=${n}', j['value'], j['ms_per_step'])
except Exception as e: print('unreadable', e)
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workload synthetic --scale 0.02 > gpurun_out/bench_syn_8gpu.json 2> gpurun_out/bench_syn_8gpu.err; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/bench_syn_8gpu.err | tail -3; cut -c1-160 gpurun_out/bench_syn_8gpu.json
