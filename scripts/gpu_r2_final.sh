#!/bin/bash
# round-2 final single-GPU call: whole GPU suite, smoke(), the default bench line (full C5 graph, e2e, CPU baseline),
# the secondary workloads, the reference arm, the ncu launch list of the default command and the DRAM traffic of the
# aggregation kernels at full size (profiles/r2_traffic.json)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2f_gpu_tests.txt
tail -6 gpurun_out/r2f_gpu_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.txt 2>&1; tail -2 gpurun_out/r2f_smoke.txt
( time timeout 1200 python bench.py > gpurun_out/r2f_full_default.json 2> gpurun_out/r2f_full_default.err ) 2> gpurun_out/r2f_full_default.time
timeout 300 python bench.py --scale 0.02 --steps 20 > gpurun_out/r2f_syn0.02.json 2> gpurun_out/r2f_syn0.02.err
timeout 300 python bench.py --workload fb15k237 --steps 50 > gpurun_out/r2f_fb.json 2> gpurun_out/r2f_fb.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_reference_arm.json 2> gpurun_out/r2f_reference_arm.err ) 2> gpurun_out/r2f_reference_arm.time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity-check > gpurun_out/r2f_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
  -k regex:"k_block_team" --launch-skip 2 --launch-count 2 --csv --page raw --log-file gpurun_out/r2f_traffic_raw.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check > gpurun_out/r2f_traffic.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2f_*.json")):
    try:
        j = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        rl = j.get("roofline_layer") or {}
        print(f, "%.1f %s %.3f ms layer_frac %s" % (j["value"], j["unit"], j.get("ms_per_step", 0), rl.get("frac")),
              {k: round(v, 3) for k, v in (j.get("stages_ms") or {}).items() if v > 0.05}, "e2e", j.get("e2e"), "cpu", j.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/r2f_full_default.time | grep real; head -5 gpurun_out/r2f_traffic_raw.csv | cut -c1-300
