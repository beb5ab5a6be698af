#!/bin/bash
# r05 final pass: tools/gpu_final.sh + ncu --set full of the attention backward + memcheck over the attention tests
TAG=${1:-r05final}
tools/gpu_final.sh $TAG
for c in attn40bwd_drop; do
  timeout 300 tools/ncu_one.sh $c attn_bwd_tc_kernel
  python tools/ncu_summary.py gpurun_out/ncu_$c.raw.csv > gpurun_out/${TAG}_ncu_full_$c.txt 2>&1
  grep -E "gpu__time_duration.sum|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_alu.avg.pct|sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active|launch__registers" gpurun_out/${TAG}_ncu_full_$c.txt
done
rm -f gpurun_out/ncu_*.ncu-rep
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention_dropout_fwd_bwd or attention_fwd_bwd" > gpurun_out/${TAG}_memcheck_attn.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck_attn.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_memcheck_attn.log | tail -4
