#!/bin/bash
# round-2 GPU pass: all parity tests (no -x: every failure is listed), smoke, one bench line.  usage: tools/gpu_r02.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_tests.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short ${PYTEST_ARGS} 2>&1 | grep -v "^E   *+\|size mismatch" | cut -c1-600 >> gpurun_out/${TAG}_tests.log
tail -40 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/print_bench.py gpurun_out/${TAG}_bench.json; tail -n 3 gpurun_out/${TAG}_bench.err | cut -c1-300
