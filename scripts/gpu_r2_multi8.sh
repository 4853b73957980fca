#!/bin/bash
# round-2 8-GPU call: world-4 equivalence tests, then the strong-scaling bench of the full graph with both halo modes
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu -k "4-" > gpurun_out/r2_parallel_tests_w4_on$N.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_parallel_tests_w4_on$N.txt
tail -3 gpurun_out/r2_parallel_tests_w4_on$N.txt
for halo in overlapped pipelined; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 8 --warmup 3 --halo $halo --no-cpu-baseline > gpurun_out/r2_scale_full_n${N}_$halo.json 2> gpurun_out/r2_scale_full_n${N}_$halo.err ) 2> gpurun_out/r2_scale_full_n${N}_$halo.time
  echo "bench $halo rc=$?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r2_scale_full_n${N}_$halo.err | tail -4; grep real gpurun_out/r2_scale_full_n${N}_$halo.time
done
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_scale_full_n${N}_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]),
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if isinstance(v, float) and v > 0.5}, "parity", (j.get("parity_check") or {}).get("max_rel_err"), "e2e", (j.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "failed", e)
PY
