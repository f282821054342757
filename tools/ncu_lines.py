#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per source line:
   python tools/ncu_lines.py both.csv [top_n]  ->  executed warp instructions, avg active threads, stall samples."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; hdr = None; lines = []
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if row[0] == "": continue
    d = dict(zip(hdr, row))
    try:
        ie = int(d["Instructions Executed"]); te = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
    except Exception:
        continue
    lines.append((cur, int(row[0]), ie, te, smp, row[1]))
tot = sum(l[2] for l in lines); tots = sum(l[4] for l in lines)
print(f"total warp-inst {tot}  samples {tots}")
byfile = collections.Counter()
for l in lines: byfile[l[0]] += l[2]
print(byfile)
for l in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{l[0]}:{l[1]:5d} inst {l[2]:>11d} ({100*l[2]/tot:4.1f}%) thr/inst {l[3]/max(l[2],1):5.1f} smp {100*l[4]/max(tots,1):4.1f}%  {l[5][:110]}")
