#!/bin/bash
# N = $1 strong-scaling bench of the default workload, peer transport, adaptive supertiles
N=${1:-8}
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_scale_full_n${N}_v2.json 2> gpurun_out/r2_scale_full_n${N}_v2.err ) 2> gpurun_out/r2_scale_full_n${N}_v2.time
echo "bench rc=$?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version\|FutureWarning\|enable_symm" gpurun_out/r2_scale_full_n${N}_v2.err | tail -8; grep real gpurun_out/r2_scale_full_n${N}_v2.time
python - <<PY
import json
f = "gpurun_out/r2_scale_full_n${N}_v2.json"
j = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]), j["config"]["parallelism"],
      {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if isinstance(v, float) and v > 0.5}, "parity", (j.get("parity_check") or {}).get("max_rel_err"))
PY
