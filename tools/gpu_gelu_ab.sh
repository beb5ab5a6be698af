#!/bin/bash
mkdir -p gpurun_out
for v in scalar packed scalar packed; do
  if [ $v = scalar ]; then export NVCC_EXTRA="-DFHB_GELU_SCALAR"; else export NVCC_EXTRA=""; fi
  python -m fithubert_b200.build --force > /dev/null 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gelu_$v.csv python tools/kernel_bench.py "S conv1" "S conv6" "T fc1" > /dev/null 2>&1
  echo "== $v"
  python - $v <<'PY'
import csv, sys
v = sys.argv[1]
lines = [l for l in open(f"gpurun_out/gelu_{v}.csv") if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if "fhb_gemm" in r["Kernel Name"]]
vals = [float(r["Metric Value"]) / 1e3 for r in rows]
for i in range(0, len(vals), 23):
    ch = vals[i + 3:i + 23]
    if ch: print(f" case {i // 23}: {sum(ch) / len(ch):7.1f} us")
PY
done
unset NVCC_EXTRA
python -m fithubert_b200.build --force > /dev/null 2>&1
