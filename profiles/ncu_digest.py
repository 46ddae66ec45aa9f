#!/usr/bin/env python
"""Digest an .ncu-rep: headline metrics per kernel + SASS hot spots (needs ncu on PATH)."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg",
        "sm__cycles_elapsed.avg ", "smsp__inst_executed.sum ", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "lts__t_bytes.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_wait_per_issue_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active", "smsp__average_warps_issue_stalled_not_selected_per_issue_active",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
name_col = hdr.index("Kernel Name")
for r in rows[2:]:
    print("=" * 100)
    print(r[name_col][:90])
    for h, u, v in zip(hdr, units, r):
        if any(h == w.strip() or h.startswith(w) for w in WANT):
            print(f"  {h:88s} {v:>18s} {u}")
