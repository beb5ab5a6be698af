#!/bin/bash
# Trace build of the library: attention_bwd_tc.cu with -DFHB_BWD_TRACE, every other object from the normal build.
set -e
cd "$(dirname "$0")/.."
python fithubert_b200/build.py > /dev/null
B=fithubert_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 -DFHB_BWD_TRACE \
  -c fithubert_b200/csrc/attention_bwd_tc.cu -o $B/attention_bwd_tc_trace.o 2>&1 | grep -v "nvcc warning" || true
nvcc -shared -o $B/libfhb_trace.so $(ls $B/*.o | grep -v "attention_bwd_tc") $B/attention_bwd_tc_trace.o -lcudart_static -ldl -lrt -lpthread 2>&1 | grep -v "nvcc warning" || true
ls -la $B/libfhb_trace.so
