"""Top stall-sample SASS lines of an ncu report: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
key = "Warp Stall Sampling (All Samples)"
tot = sum(int(r[key] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
for i, r in enumerate(rows):
    r["_i"] = i
for r in sorted(rows, key=lambda r: -int(r[key] or 0))[:n]:
    print(f"{r['_i']:5d} {100*int(r[key] or 0)/max(tot,1):5.1f}%  exec {r['Instructions Executed']:>9s}  {r['Source'].strip()[:110]}")
