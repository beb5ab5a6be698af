#!/bin/bash
# One GPU round without ncu: parity tests + bench line.  usage: tools/gpu_ab.sh TAG [extra bench args]
TAG=${1:-ab}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_tests.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^E   *+" | tail -40 >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
