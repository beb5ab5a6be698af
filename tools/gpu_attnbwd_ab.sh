#!/bin/bash
# A/B of the attention kernels: parity tests on the new build, then isolated timings of both builds (interleaved).
# usage: tools/gpu_attnbwd_ab.sh TAG   (baseline library: fithubert_b200/build/libfhb_base.so)
TAG=${1:-attnbwd}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention or dropout" 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for rep in 1 2 3; do
  echo "new"; timeout 200 python tools/kernel_bench.py "attn bwd" "attn fwd"
  echo "base"; FHB_LIB=$PWD/fithubert_b200/build/libfhb_base.so timeout 200 python tools/kernel_bench.py "attn bwd" "attn fwd"
done 2>&1 | tee gpurun_out/${TAG}_ab.txt
