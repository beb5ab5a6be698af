#!/bin/bash
# Run each check case in its own process with a timeout; log to gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/${1:-suite}.log
shift
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
for c in "$@"; do
  echo "=== $c" >> $LOG
  timeout 180 python tools/gemm_check.py $c >> $LOG 2>&1
  echo "exit $?" >> $LOG
done
tail -100 $LOG
