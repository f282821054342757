#!/usr/bin/env python
"""Hot SASS region of `ncu --page source --print-source sass --csv` with executed counts and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 3e6
hdr = rows[1]
ia, isrc, ie, it, ism = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
out = []
for r in rows[2:]:
    try: e = int(r[ie])
    except ValueError: continue
    st = sorted(((int(r[i]), h[6:]) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
    out.append((r[isrc].strip(), e, int(r[it]), int(r[ism]), st))
hot = [i for i, o in enumerate(out) if o[1] > thr]
tot = sum(o[3] for o in out)
print("total samples", tot)
for i in range(hot[0] - 2, hot[-1] + 3):
    o = out[i]
    print(f"{i:5d} {o[1]:>10d} {o[2]/max(o[1],1):5.1f} {o[3]:5d} {' '.join(f'{n}:{c}' for c, n in o[4]):28s} {o[0]}")
