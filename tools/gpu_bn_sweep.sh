#!/bin/bash
# ncu-timed sweep of the forced GEMM tile width over the student-sized GEMMs
mkdir -p gpurun_out
: > gpurun_out/bn_sweep.log
for bn in 0 64 128 192 256; do
  FHB_GEMM_BN=$bn ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bn_$bn.csv \
    python tools/kernel_bench.py "S qkv" "S out" "S fc1 plain" "S dgrad fc" "S dgrad qkv" "S wgrad fc" "S wgrad qkv" "T out" > /dev/null 2>&1
  echo "== bn $bn" >> gpurun_out/bn_sweep.log
  python - "$bn" >> gpurun_out/bn_sweep.log <<'PY'
import csv, sys, collections
bn = sys.argv[1]
lines = [l for l in open(f"gpurun_out/bn_{bn}.csv") if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if "fhb_gemm" in r["Kernel Name"]]
# kernel_bench runs every case 23 times (3 warm-up + 20): group consecutive launches
vals = [float(r["Metric Value"]) / 1e3 for r in rows]
per = 23
for i in range(0, len(vals), per):
    chunk = vals[i + 3:i + per]
    if chunk:
        print(f"case {i // per}: {sum(chunk) / len(chunk):7.1f} us  grid {rows[i]['Grid Size']}")
PY
done
cat gpurun_out/bn_sweep.log
