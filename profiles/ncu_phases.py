#!/usr/bin/env python
"""Executed warp instructions / stall samples of dvs_tile_fast.cuh per phase (line ranges given on the command line).
usage: python profiles/ncu_phases.py <report.ncu-rep> name:lo-hi [name:lo-hi ...]; other files are grouped by file"""
import csv, io, os, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
phases = []
for a in sys.argv[2:]:
    n, r = a.split(":"); lo, hi = r.split("-"); phases.append((n, int(lo), int(hi)))
import re
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
file_phases = {}  # file name -> [(name, lo, hi)]
if phases:
    file_phases["dvs_tile_fast.cuh"] = phases
else:  # the numbered phase comments ("// ---- 2. ...") of the kernel headers
    for fn in ("dvs_tile_fast.cuh", "dvs_tile_ranked.cuh"):
        try: src = open(os.path.join(here, "pydvs_b200", "csrc", fn)).read().splitlines()
        except OSError: continue
        marks = [(i + 1, re.search(r"// ---- ([1-4][ab]?\. [^-]*)", l).group(1).strip()) for i, l in enumerate(src) if re.search(r"// ---- [1-4][ab]?\. ", l)]
        end = max(i + 1 for i, l in enumerate(src) if "write_counters(A, s_idx, t_idx, hist, nh" in l)
        file_phases[fn] = [(n[:44], ln, marks[k + 1][0] - 1 if k + 1 < len(marks) else end) for k, (ln, n) in enumerate(marks)]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr = None, None
ins, smp = defaultdict(int), defaultdict(int)
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1])
    elif r[0] == "Line No": hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try: i, s = int(d["Instructions Executed"]), int(d["# Samples"])
        except Exception: continue
        key = cur
        if cur in file_phases:
            key = cur + ":other"
            for n, lo, hi in file_phases[cur]:
                if lo <= int(r[0]) <= hi: key = n; break
        ins[key] += i; smp[key] += s
ti, ts = sum(ins.values()) or 1, sum(smp.values()) or 1
for k in sorted(ins, key=lambda k: -ins[k]):
    print("%-46s %6.1f%% inst (%d)  %5.1f%% samples" % (k, 100.0 * ins[k] / ti, ins[k], 100.0 * smp[k] / ts))
