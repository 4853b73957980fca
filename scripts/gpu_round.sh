#!/bin/bash
set -x
mkdir -p gpurun_out
for nv in 2 1; do for nf in 0 1; do
  if [ $nf = 1 ]; then export RGCN_NO_FUSE_DW=1; else unset RGCN_NO_FUSE_DW; fi
  RGCN_REL_NV=$nv timeout 600 python bench.py --workload synthetic --scale 0.02 --no-cpu-baseline --steps 5 > gpurun_out/bench_syn_nv${nv}_nf${nf}.json 2> gpurun_out/bench_syn.err; echo "bench syn rc=$?"
  tail -3 gpurun_out/bench_syn.err; python scripts/show_bench.py gpurun_out/bench_syn_nv${nv}_nf${nf}.json
done; done
