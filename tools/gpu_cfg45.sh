#!/bin/bash
# New BASELINE configs (cfg-4 expert forward, cfg-5 FitW2V2 30 s): parity tests, then the bench lines.
TAG=${1:-r01z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_tests.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -s -k "cfg4 or cfg5" 2>&1 | grep -v "^E   *+" | tail -40 >> gpurun_out/${TAG}_tests.log
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | grep -v "^E   *+" | tail -40 >> gpurun_out/${TAG}_tests.log
tail -30 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
cat gpurun_out/${TAG}_bench_cfg4.json; tail -n 3 gpurun_out/${TAG}_bench_cfg4.err
