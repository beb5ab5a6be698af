#!/bin/bash
# interleaved A/B of programmatic dependent launch: ms_per_step for off/on/off/on/off/on
mkdir -p gpurun_out
: > gpurun_out/pdl_ab.log
for i in 1 2 3; do
  for mode in off on; do
    if [ $mode = on ]; then export FHB_PDL=1; else unset FHB_PDL; fi
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', d['ms_per_step'], d['e2e']['value'], d['clocks'])" >> gpurun_out/pdl_ab.log
  done
done
cat gpurun_out/pdl_ab.log
