#!/bin/bash
# A/B of the chunked host->device copy (FHB_H2D_CHUNKS) on the e2e number; parity tests of the host-buffer paths first.
TAG=${1:-r01zz_h2d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "full_size or fused_step or fit_loop or cfg5 or cosine" 2>&1 | grep -v "^E   *+" | tail -15 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
LOG=gpurun_out/${TAG}_ab.log
: > $LOG
for rep in 1 2; do
  for v in 1 2 4 8; do
    echo "=== FHB_H2D_CHUNKS=$v rep $rep" >> $LOG
    FHB_H2D_CHUNKS=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-student-fwd 2>gpurun_out/${TAG}_err.log \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step %.3f value %.0f e2e %.0f loss %.6f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['loss']))" >> $LOG 2>&1
    tail -n 2 gpurun_out/${TAG}_err.log | cut -c1-300 >> $LOG
  done
done
cat $LOG
