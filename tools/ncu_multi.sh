#!/bin/bash
# usage: tools/ncu_multi.sh "<case> <kernel-regex>" ...   -> gpurun_out/ncu_<case>.{ncu-rep,raw.csv,txt}
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec
  CASE=$1; KRE=$2
  ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip 2 --launch-count 1 \
    -o gpurun_out/ncu_$CASE -f python tools/one_kernel.py $CASE > gpurun_out/ncu_$CASE.log 2>&1
  ncu -i gpurun_out/ncu_$CASE.ncu-rep --page raw --csv > gpurun_out/ncu_$CASE.raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/ncu_$CASE.raw.csv > gpurun_out/ncu_$CASE.txt 2>&1
  tail -2 gpurun_out/ncu_$CASE.log
done
