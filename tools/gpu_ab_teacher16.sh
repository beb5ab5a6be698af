#!/bin/bash
# A/B: teacher residual stream fp32 (default) vs fp16: parity of the full-geometry test + step time, interleaved
mkdir -p gpurun_out
for v in 1 0 1 0; do
  FHB_TEACHER_STREAM32=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-student-fwd > gpurun_out/ab_t32_$v.json 2>/dev/null
  echo "stream32=$v $(python tools/print_bench.py gpurun_out/ab_t32_$v.json)"
done
FHB_TEACHER_STREAM32=0 timeout 600 python -m pytest tests -m gpu -q -s -k "full_geometry or cfg5" 2>&1 | grep -v "^E   *+" | tail -8
sort -r gpurun_out/parity_full_geometry_activations.txt | head -5
sort -r gpurun_out/parity_full_geometry_grads.txt | head -3
