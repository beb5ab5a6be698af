"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG /
UTMAREDG = TMA load / store / reduce, LDTM / STTM = tcgen05.ld / st, HMMA = mma.sync) in libfhb_sm100a.so.
Runs without a GPU: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fithubert_b200", "libfhb_sm100a.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "HMMA", "F2FP", "MUFU.EX2")
cur, tab = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        tab[cur] = collections.Counter()
        continue
    if cur:
        for n in names:
            if re.search(r"\b" + re.escape(n), line):
                tab[cur][n] += 1
print(f"{'kernel':64s} " + " ".join(f"{n:>9s}" for n in names))
tot = collections.Counter()
for k, c in tab.items():
    if sum(c.values()):
        print(f"{k[:64]:64s} " + " ".join(f"{c[n]:9d}" for n in names))
        tot.update(c)
print(f"{'TOTAL':64s} " + " ".join(f"{tot[n]:9d}" for n in names))
