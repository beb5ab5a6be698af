"""Is the step GPU-bound between its short kernels?  Reads the log the trace build of the library keeps of every GEMM
launch (CTA 0 start / end, GPU nanoseconds) over 3 training steps queued back to back:
  tools/build_gemm_trace_lib.sh
  FHB_LIB=$PWD/fithubert_b200/build/libfhb_gemmtrace.so FHB_GEMM_STEPLOG=gpurun_out/steplog.txt python bench.py --profile
  python tools/gemm_steplog.py gpurun_out/steplog.txt
Prints per GEMM shape: launches per step, CTA-0 lifetime, and the gap to the PREVIOUS GEMM's end when that GEMM is the
directly preceding kernel of the stream in program order is unknown to this log - so the gap column is an upper bound of
the launch gap (it contains whatever non-GEMM kernels ran in between)."""
import sys
from collections import defaultdict

rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1])]
rows.sort()
n = len(rows)
per_step = n // 3
last = rows[2 * per_step:]          # the third step
t0 = last[0][0]
span = (last[-1][1] - t0) / 1e3
busy = sum(e - s for s, e, *_ in last) / 1e3
mhz = sorted(1e3 * r[6] / (r[1] - r[0]) for r in last if len(r) > 6 and r[1] - r[0] > 20000)
if mhz:
    print(f"SM clock during the GEMMs of the step (clock64 / globaltimer over launches longer than 20 us): min {mhz[0]:.0f} median {mhz[len(mhz) // 2]:.0f} max {mhz[-1]:.0f} MHz")
print(f"{n} GEMM launches logged, {per_step} per step; third step: first GEMM start -> last GEMM end {span:.1f} us, sum of CTA-0 lifetimes {busy:.1f} us")
agg = defaultdict(lambda: [0, 0.0, 0.0, 1e9])
prev_end = None
for s, e, m, nn, k, fl, *_ in last:
    a = agg[(m, nn, k, fl)]
    a[0] += 1
    a[1] += (e - s) / 1e3
    if prev_end is not None:
        g = (s - prev_end) / 1e3
        a[2] += g
        a[3] = min(a[3], g)
    prev_end = e
print(f"{'m':>9} {'n':>6} {'k':>8} {'flags':>8} {'cnt':>4} {'life us':>9} {'gap-to-prev avg':>16} {'min':>7}")
for (m, nn, k, fl), (c, life, gap, gmin) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{m:9d} {nn:6d} {k:8d} {fl:#8x} {c:4d} {life / c:9.1f} {gap / c:16.1f} {gmin:7.1f}")
