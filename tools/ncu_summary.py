"""Print the metrics that matter from an `ncu --page raw --csv` export."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput",
        "l1tex__throughput", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak",
        "smsp__average_warp", "smsp__warp_issue_stalled", "smsp__average_warps_issue_stalled", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu",
        "smsp__inst_executed_pipe", "sm__pipe_", "l1tex__data_bank_conflicts", "smsp__pcsamp", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "lts__t_sector_hit_rate", "sm__ctas_launched"]
path = sys.argv[1]
pat = sys.argv[2:] or KEYS
rows = list(csv.reader(open(path)))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if any(k in h for k in pat):
        print(f"{h:90s} {v:>18s} {u}")
