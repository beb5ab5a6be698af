#!/bin/bash
# Trace build of the library: gemm_sm100.cu with -DFHB_GEMM_TRACE, every other object from the normal build
# (-> fithubert_b200/build/libfhb_gemmtrace.so; run tools/gemm_trace.py with FHB_LIB pointing at it).
set -e
cd "$(dirname "$0")/.."
python fithubert_b200/build.py > /dev/null
B=fithubert_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 -DFHB_GEMM_TRACE \
  -c fithubert_b200/csrc/gemm_sm100.cu -o $B/gemm_sm100_trace.o 2>&1 | grep -v "nvcc warning" || true
nvcc -shared -o $B/libfhb_gemmtrace.so $(ls $B/*.o | grep -v "gemm_sm100\|_trace.o") $B/gemm_sm100_trace.o -lcudart_static -ldl -lrt -lpthread 2>&1 | grep -v "nvcc warning" || true
ls -la $B/libfhb_gemmtrace.so
