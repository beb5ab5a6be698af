#!/bin/bash
# interleaved A/B of programmatic dependent launch: ms_per_step for mode 0 / 2 / 1, three rounds
mkdir -p gpurun_out
: > gpurun_out/pdl_ab.log
for i in 1 2 3; do
  for mode in 0 2 1; do
    if [ $mode = 0 ]; then unset FHB_PDL; else export FHB_PDL=$mode; fi
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode $mode', d['ms_per_step'], d['e2e']['value'], d['clocks'])" >> gpurun_out/pdl_ab.log
  done
done
cat gpurun_out/pdl_ab.log
