"""Executed-instruction mix by opcode from an ncu report: python tools/ncu_mix.py report.ncu-rep"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
mix = collections.Counter()
for r in rows:
    src = r["Source"].strip()
    if src.startswith("@"):
        src = src.split(None, 1)[1]
    op = src.split()[0].split(".")[0]
    mix[op] += int(r["Instructions Executed"] or 0)
tot = sum(mix.values())
print("total warp instructions", tot)
for op, n in mix.most_common(30):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%")
