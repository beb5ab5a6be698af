#!/bin/bash
# quick check: GPU parity tests, then the bench line twice (e2e + value).  usage: tools/gpu_quick.sh TAG
TAG=${1:-quick}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x  2>&1 | grep -v "^E   *+" | tail -15 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-student-fwd > gpurun_out/${TAG}_bench$rep.json 2> gpurun_out/${TAG}_bench$rep.err
  python tools/print_bench.py gpurun_out/${TAG}_bench$rep.json
  tail -n 2 gpurun_out/${TAG}_bench$rep.err | cut -c1-300
done
