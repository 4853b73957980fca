#!/bin/bash
# round-2 multi-GPU call (N = $1): peer-mapped halo transport -- equivalence tests, then the strong-scaling bench
N=${1:-2}
MODES=${2:-"peer nccl"}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x -k "${3:-2 and (peer or device or overlapped)}" > gpurun_out/r2_peer_tests_n$N.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_peer_tests_n$N.txt
tail -15 gpurun_out/r2_peer_tests_n$N.txt
for tr in $MODES; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 6 --warmup 3 --transport $tr --no-cpu-baseline --no-e2e > gpurun_out/r2_scale_full_n${N}_$tr.json 2> gpurun_out/r2_scale_full_n${N}_$tr.err ) 2> gpurun_out/r2_scale_full_n${N}_$tr.time
  echo "bench $tr rc=$?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r2_scale_full_n${N}_$tr.err | tail -12; grep real gpurun_out/r2_scale_full_n${N}_$tr.time
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_scale_full_n${N}_*.json")):
    try:
        j = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "%.1f M-edges/s %.3f ms layer_frac %.3f" % (j["value"], j["ms_per_step"], j["roofline_layer"]["frac"]), j["config"]["parallelism"],
              {k: round(v, 3) for k, v in (j["stages_ms"] or {}).items() if isinstance(v, float) and v > 0.5}, "parity", (j.get("parity_check") or {}).get("max_rel_err"), (j.get("parity_check") or {}).get("what"))
    except Exception as e:
        print(f, "failed", e)
PY
