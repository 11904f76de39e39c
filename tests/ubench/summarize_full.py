#!/usr/bin/env python
"""Summarise `ncu -i <rep> --page raw --csv` of an `--set full` capture: one row per launch, the columns the roofline
argument needs (duration, DRAM bytes, DRAM / L1 / issue utilisation, shared-memory wavefronts and bank conflicts, lanes
active per instruction, the dominant stall reasons).
usage: ncu -i gpurun_out/<tag>.ncu-rep --page raw --csv > /tmp/full.csv; python tests/ubench/summarize_full.py /tmp/full.csv "<header comment>" > profiles/<tag>_summary.csv"""
import csv
import re
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__cycles_elapsed.max", "launch__cluster_size"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) for c in COLS if c in hdr]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in rows[2:]:
    out = [r[i] for i in idx]
    out[0] = re.sub(r"\(.*", "", out[0]).replace("void ", "").replace("gb::", "")
    w.writerow(out)
