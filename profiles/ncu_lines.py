#!/usr/bin/env python
"""Per-source-line totals of an .ncu-rep captured with --import-source on (kernels built with -lineinfo).
usage: python profiles/ncu_lines.py <report.ncu-rep> [top-N]   -> executed warp instructions and stall samples per line"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1])
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            out.append((int(d["Instructions Executed"]), int(d["# Samples"]), cur, int(r[0]), r[1].strip()))
        except Exception:
            pass
tot_i = sum(o[0] for o in out) or 1
tot_s = sum(o[1] for o in out) or 1
print("total executed warp instructions %d, samples %d" % (tot_i, tot_s))
for ins, smp, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * ins / tot_i, 100.0 * smp / tot_s, f, ln, src[:110]))
