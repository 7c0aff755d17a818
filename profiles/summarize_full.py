#!/usr/bin/env python
"""Headline counters of `ncu --set full` reports -> one CSV (header, units, one row per
kernel), the file bench.py reads `roofline.traffic` from.

usage: summarize_full.py out.csv report1.ncu-rep [report2.ncu-rep ...]
Runs here (no GPU): ncu -i <rep> --page raw --csv.
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg"]

# ncu picks a unit per report and metric: bring bytes to "byte" and times to "us"
SCALE = {"byte": ("byte", 1.0), "Kbyte": ("byte", 1e3), "Mbyte": ("byte", 1e6), "Gbyte": ("byte", 1e9),
         "ns": ("us", 1e-3), "us": ("us", 1.0), "ms": ("us", 1e3), "s": ("us", 1e6)}

out, reps = sys.argv[1], sys.argv[2:]
rows, units_row = [], None
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    table = list(csv.reader(io.StringIO(raw)))
    hdr, units = table[0], table[1]
    for r in table[2:]:
        vals, us = [], []
        for k in KEYS:
            v, u = r[hdr.index(k)], units[hdr.index(k)]
            if u in SCALE:
                v = repr(float(v.replace(",", "")) * SCALE[u][1])
                u = SCALE[u][0]
            vals.append(v)
            us.append(u)
        rows.append([r[hdr.index("Kernel Name")]] + vals)
        units_row = [""] + us
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["Kernel Name"] + KEYS)
    w.writerow(units_row)
    w.writerows(rows)
print("wrote", out, len(rows), "kernels")
