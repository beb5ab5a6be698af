#!/bin/bash
# A/B of FHB_STREAMS (bit 1: teacher fwd || student fwd; bit 2: layer wgrads on a side stream; bit 4: conv wgrads too)
# and FHB_HEAD_COMPOSE (folded projection heads), interleaved twice; parity tests of the new modes first.
TAG=${1:-r01z_streams}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "folded or side_streams or fixture or fused_step or full_size or cfg5" 2>&1 | grep -v "^E   *+" | tail -25 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
LOG=gpurun_out/${TAG}_ab.log
: > $LOG
for rep in 1 2; do
  for v in "0 0" "0 1" "1 1" "2 1" "3 1" "7 1"; do
    set -- $v
    echo "=== FHB_STREAMS=$1 FHB_HEAD_COMPOSE=$2 rep $rep" >> $LOG
    FHB_STREAMS=$1 FHB_HEAD_COMPOSE=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-student-fwd 2>gpurun_out/${TAG}_err.log \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step %.3f value %.0f e2e %.0f loss %.6f gemm_ms %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['loss'], d['roofline']['gemm_ms_per_step']))" >> $LOG 2>&1
    tail -n 2 gpurun_out/${TAG}_err.log | cut -c1-300 >> $LOG
  done
done
cat $LOG
