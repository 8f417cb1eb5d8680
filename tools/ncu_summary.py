#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into the few numbers the design notes track."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum"]
for r in rows[2:3 + int(sys.argv[2]) if len(sys.argv) > 2 else 3]:
    print("----")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(src.splitlines()))
rows = []
for r in allrows[2:]:
    if len(r) < 6 or not r[0].startswith("0x"):
        break
    rows.append(r)
ti = sum(int(r[5]) for r in rows)
ts = sum(int(r[2]) for r in rows)
print("SASS lines", len(rows), "warp inst", ti, "samples", ts)
h, hs = collections.Counter(), collections.Counter()
for r in rows:
    t = r[1].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    h[op] += int(r[5]); hs[op] += int(r[2])
for op, c in h.most_common(16):
    print(f"  {op:10s} inst {c / ti * 100:5.1f}%  stall-samples {hs[op] / max(ts, 1) * 100:5.1f}%")
print("top stall lines:")
for r in sorted(rows, key=lambda r: -int(r[2]))[:14]:
    print("  ", r[2], r[5], r[1].strip()[:90])
