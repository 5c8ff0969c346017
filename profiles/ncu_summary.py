#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics + SASS-level stall / opcode mix.
usage: python profiles/ncu_summary.py <report.ncu-rep> [kernel-index]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__block_size",
        "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_elapsed.max", "launch__waves_per_multiprocessor"]
for r in rows[2:][which:which + 1]:
    d = dict(zip(hdr, r))
    for w in want:
        if w in d:
            print("%-70s %s %s" % (w, d[w], units[hdr.index(w)]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                      text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
b0 = idx[which]
b1 = idx[which + 1] if which + 1 < len(idx) else len(rows)
hdr = rows[b0 + 1]
blk = rows[b0 + 2:b1]
ci = hdr.index("Instructions Executed")
tot = sum(int(r[ci]) for r in blk if len(r) > ci and r[ci].isdigit())
print("static SASS instructions: %d, executed warp instructions: %d" % (len(blk), tot))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = Counter()
for r in blk:
    for i in stall_cols:
        try:
            agg[hdr[i]] += int(r[i])
        except Exception:
            pass
s = sum(agg.values()) or 1
print("stall samples:", ", ".join("%s %.1f%%" % (k, 100 * v / s) for k, v in agg.most_common(8)))
ops = Counter()
for r in blk:
    try:
        toks = r[1].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += int(r[ci])
    except Exception:
        pass
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(16)))
