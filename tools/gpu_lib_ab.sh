#!/bin/bash
# Full GPU parity suite on the in-tree build, then interleaved bench pairs: in-tree build vs another build of the library.
# usage: tools/gpu_lib_ab.sh TAG BASE_LIB [pairs]
TAG=${1:-libab}; BASE=$2; PAIRS=${3:-3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^E   *+" | tail -6 > gpurun_out/${TAG}_tests.log
tail -2 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_ab.txt
for rep in $(seq 1 $PAIRS); do
  for which in new base; do
    if [ $which = base ]; then export FHB_LIB=$PWD/$BASE; else unset FHB_LIB; fi
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-student-fwd > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err
    python - $which gpurun_out/${TAG}_b.json >> gpurun_out/${TAG}_ab.txt <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
g = d["roofline"]["groups"]
print("%-4s ms_per_step %.3f value %.0f e2e %.0f | gemm ms %.3f conv_stacks %.3f student_and_rest %.3f teacher %.3f | frac %.3f" % (
    sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["gemm_ms_per_step"], g["conv_stacks"]["ms_per_step"],
    g["student_and_rest"]["ms_per_step"], g["teacher_encoder"]["ms_per_step"], d["roofline"]["frac"]))
PY
  done
done
cat gpurun_out/${TAG}_ab.txt
