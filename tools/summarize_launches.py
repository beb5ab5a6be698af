"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((int(r["ID"]), r["Kernel Name"], v))
    if skip == "step":
        # keep the last full step: from just after the second-to-last adamw launch to the last one (inclusive)
        idx = [i for i, r in enumerate(rows) if "adamw" in r[1]]
        if len(idx) >= 2:
            rows = rows[idx[-2] + 1:idx[-1] + 1]
    else:
        rows = rows[int(skip):]
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for _, k, v in rows:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"<unnamed>::", "", k)
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    print(f"launches {len(rows)}  total {total / 1e3:.3f} ms (serialised, cold-cache)")
    print(f"{'kernel':60s} {'count':>6s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k[:60]:60s} {cnt[k]:6d} {v / 1e3:9.3f} {100 * v / total:6.1f}% {v / cnt[k]:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 0)
