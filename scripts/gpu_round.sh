#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --triples-npz .scratch/fb15k237_train.npz --no-cpu-baseline > gpurun_out/bench_real_fb15k237.json 2> gpurun_out/bench_real.err; tail -3 gpurun_out/bench_real.err; python scripts/show_bench.py gpurun_out/bench_real_fb15k237.json
timeout 600 python bench.py --triples-npz .scratch/fb15k_train.npz --no-cpu-baseline --steps 100 > gpurun_out/bench_real_fb15k.json 2> gpurun_out/bench_real.err; tail -3 gpurun_out/bench_real.err; python scripts/show_bench.py gpurun_out/bench_real_fb15k.json
RGCN_BLOCK_ALGO=0 timeout 600 python bench.py --triples-npz .scratch/fb15k237_train.npz --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/bench_real_fb15k237_dstmajor.json 2> gpurun_out/bench_real.err; tail -3 gpurun_out/bench_real.err; python scripts/show_bench.py gpurun_out/bench_real_fb15k237_dstmajor.json
