#!/bin/bash
# Round-final pass on one B200: all GPU tests, smoke, the default bench line, the ncu launch list of one step, the per-shape
# GEMM table.  usage: tools/gpu_final.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_tests.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s --tb=short 2>&1 | grep -v "^E   *+\|size mismatch" | cut -c1-600 >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/print_bench.py gpurun_out/${TAG}_bench.json; tail -n 2 gpurun_out/${TAG}_bench.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-400 gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --profile > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv step > gpurun_out/${TAG}_launch_summary.txt 2>&1
head -12 gpurun_out/${TAG}_launch_summary.txt
timeout 600 python tools/gemm_table.py > gpurun_out/${TAG}_gemm_table.txt 2>&1; head -3 gpurun_out/${TAG}_gemm_table.txt
