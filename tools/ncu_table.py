#!/usr/bin/env python
"""Markdown table (one row per kernel launch) from `ncu -i X.ncu-rep --page raw --csv`.
    python tools/ncu_table.py raw.csv > table.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = lambda r, k: r[hdr.index(k)] if k in hdr else ""
f = lambda x: float(x.replace(",", "")) if x else float("nan")
print("| # | kernel | time µs | DRAM read MB | DRAM write MB | DRAM % of peak | issue-active % | warps active % | regs | warp-inst M |")
print("|---|---|---|---|---|---|---|---|---|---|")
for i, r in enumerate(data):
    name = col(r, "Kernel Name").split("(")[0].replace("void ", "").replace("cugs::", "")
    print(f"| {i} | `{name}` | {f(col(r, 'gpu__time_duration.sum')):.1f} | {f(col(r, 'dram__bytes_read.sum')):.1f} | "
          f"{f(col(r, 'dram__bytes_write.sum')):.1f} | {f(col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
          f"{f(col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
          f"{f(col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | {col(r, 'launch__registers_per_thread')} | "
          f"{f(col(r, 'smsp__inst_executed.sum')) / 1e6:.1f} |")
