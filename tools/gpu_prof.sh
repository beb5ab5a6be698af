#!/bin/bash
# launch list of one bench step under ncu + per-shape GEMM table.  usage: tools/gpu_prof.sh TAG
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --profile > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv step > gpurun_out/${TAG}_launch_summary.txt 2>&1
head -60 gpurun_out/${TAG}_launch_summary.txt
timeout 600 python tools/gemm_table.py > gpurun_out/${TAG}_gemm_table.txt 2>&1; tail -70 gpurun_out/${TAG}_gemm_table.txt
