#!/usr/bin/env python
"""Summarise an ncu report (read here, without a GPU): python tools/ncu_summary.py rep.ncu-rep [pattern ...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
    "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_", "smsp__issue_active.avg.pct", "bank_conflicts", "wavefronts_mem_shared",
    "thread_inst_executed_per_inst", "sm__pipe_", "smsp__warps_eligible", "smsp__warps_active", "smsp__average_warp", "smsp__pcsamp_warps_issue_stalled",
    "l1tex__t_sectors_pipe_lsu_mem_global", "lts__t_sectors_op", "achieved_occupancy", "smsp__inst_executed_op_shared", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            print(f"  {h:100s} {r[i]:>18s} {units[i]}")
