#!/bin/bash
# A/B of the attention backward kernel: parity tests on the new build, then isolated timings of both builds (interleaved).
# usage: tools/gpu_attnbwd_ab.sh TAG   (baseline library: fithubert_b200/build/libfhb_base.so)
TAG=${1:-attnbwd}
mkdir -p gpurun_out
true
true
for rep in 1 2 3; do
  echo "new"; timeout 200 python tools/kernel_bench.py "attn bwd" "attn fwd d=40"
  echo "base"; FHB_LIB=$PWD/fithubert_b200/build/libfhb_base.so timeout 200 python tools/kernel_bench.py "attn bwd" "attn fwd d=40"
done 2>&1 | tee gpurun_out/${TAG}_ab.txt
