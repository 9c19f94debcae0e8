#!/usr/bin/env python3
"""SASS-level hot spots of one kernel from an .ncu-rep: instructions sorted by stall samples, with neighbours.
Usage: tools/ncu_sass.py rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(out)) if r and r[0].startswith("0x")]
tot = sum(int(r[4]) for r in rows) or 1
ti = sum(int(r[5]) for r in rows) or 1
print("%d sass instructions, %d samples, %d warp-instr executed" % (len(rows), tot, ti))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][4]))[:ntop]
for i in order:
    r = rows[i]
    print("%5d  %5.2f%% samp  %5.2f%% inst  %s" % (i, 100 * int(r[4]) / tot, 100 * int(r[5]) / ti, r[1].strip()[:90]))
