#!/bin/bash
for pp in 0 1 2 3; do
  NVCC_EXTRA="-DFHB_ATTN_POLY_PAIRS=$pp" python -m fithubert_b200.build --force > /dev/null 2>&1
  echo "== poly pairs $pp"
  python tools/kernel_bench.py "attn fwd"
  python -m pytest tests -m gpu -q -x -k "attention" 2>&1 | tail -1
done
