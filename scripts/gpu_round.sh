#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_fb.json 2> gpurun_out/bench_fb.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_fb.err; python scripts/show_bench.py gpurun_out/bench_fb.json
timeout 600 python bench.py --workload synthetic --scale 0.02 --no-cpu-baseline --steps 5 > gpurun_out/bench_syn.json 2> gpurun_out/bench_syn.err; echo "bench syn rc=$?"
tail -3 gpurun_out/bench_syn.err; python scripts/show_bench.py gpurun_out/bench_syn.json
# launch list + full capture of the hot kernels of the default bench
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_block_rel|k_block_dw|k_gemm_tf32x3' -s 12 -c 5 -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
