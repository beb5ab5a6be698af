#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/pdl_limit.log
for i in 1 2; do
  for lim in 7104 30000 150000 0; do
    FHB_PDL_GEMM_LIMIT=$lim python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('limit $lim', d['ms_per_step'])" >> gpurun_out/pdl_limit.log
  done
done
cat gpurun_out/pdl_limit.log
