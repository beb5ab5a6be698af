#!/bin/bash
# One GPU round: parity tests, bench line, ncu launch list of the same step.  usage: tools/gpu_round.sh TAG [steps]
TAG=${1:-run}
STEPS=${2:-10}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_tests.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^E   *+" | tail -40 >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps $STEPS --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --profile > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv step > gpurun_out/${TAG}_launch_summary.txt 2>&1
head -45 gpurun_out/${TAG}_launch_summary.txt
