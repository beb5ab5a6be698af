"""DRAM traffic of the fhb_gemm launches of ONE bench step from an ncu csv
(--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:fhb_gemm_kernel over `bench.py --profile`,
which runs two steps: the second one is kept).  Writes the per-launch average that bench.py reports as roofline.traffic.
usage: python tools/gemm_traffic.py gpurun_out/X.csv profiles/gemm_traffic.json"""
import csv
import json
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path, out):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = defaultdict(dict)
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1.0)
        per[int(r["ID"])][r["Metric Name"]] = v
        per[int(r["ID"])]["name"] = r["Kernel Name"].split("(")[0]
    ids = sorted(per)
    half = len(ids) // 2
    ids = ids[half:]  # the second step
    rd = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in ids)
    wr = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in ids)
    us = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in ids)
    res = {"launches_per_step": len(ids), "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
           "traffic_bytes_per_launch": (rd + wr) / max(1, len(ids)), "gemm_us_per_step_under_ncu": us,
           "dram_gbs_over_gemm_time": (rd + wr) / max(us, 1e-9) / 1e3,
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:fhb_gemm_kernel over bench.py --profile "
                     "(second of two cfg-2 steps; serialised, cold-cache launches)"}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
