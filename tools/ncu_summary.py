#!/usr/bin/env python
"""Markdown summary of `ncu -i X.ncu-rep --page raw --csv` for the metrics the roofline discussion uses.
usage: ncu_summary.py raw.csv > profiles/xxx.md"""
import csv, sys
WANT = ["gpu__time_duration.sum", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
h, units, body = rows[0], rows[1], rows[2:]
ci = {x: i for i, x in enumerate(h)}
names = [r[ci["Kernel Name"]].split("(")[0].replace("void ", "") for r in body]
print("| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |")
print("|---|---|" + "---|" * len(names))
for w in WANT:
    if w in ci:
        print(f"| {w} | {units[ci[w]]} | " + " | ".join(r[ci[w]] for r in body) + " |")
