#!/bin/bash
# round-2 multi-GPU call (N = $1): sharded == single-GPU tests, then the strong-scaling bench of the default workload
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x > gpurun_out/r2_parallel_tests_n$N.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_parallel_tests_n$N.txt
tail -4 gpurun_out/r2_parallel_tests_n$N.txt
for sc in 0.1 full; do
  extra=""; [ "$sc" != "full" ] && extra="--scale $sc"
  ( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 $extra > gpurun_out/r2_scale_${sc}_n$N.json 2> gpurun_out/r2_scale_${sc}_n$N.err ) 2> gpurun_out/r2_scale_${sc}_n$N.time
  echo "bench $sc rc=$?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r2_scale_${sc}_n$N.err | tail -5; cat gpurun_out/r2_scale_${sc}_n$N.time | grep real
done
nvidia-smi --query-gpu=index,memory.used --format=csv
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_scale_*_n$N.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if isinstance(v, float) and v > 0.5}, "parity", j.get("parity_check"), "e2e", j.get("e2e"))
    except Exception as e:
        print(f, "failed", e)
PY
