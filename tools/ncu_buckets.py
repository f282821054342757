#!/usr/bin/env python
"""Per-function and per-code-region totals of `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`:
executed warp instructions and stall samples of the step kernel split into rays / contact world / dynamics / rules / kernel.
   python tools/ncu_buckets.py src.csv [steps_in_launch] [ctas]"""
import csv, sys, collections
path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ctas = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cur = fn = None; hdr = None
fun = collections.defaultdict(lambda: [0, 0, 0]); reg = collections.defaultdict(lambda: [0, 0, 0])
def region(f, line):
    if f == "ncg_b2.cuh":
        return "b2 geometry (collide / GJK / TOI root finder)"
    if f == "ncg_car.cuh":
        # (line ranges of csrc/ncg_car.cuh as of round 2's final source: step_with_contacts 660, body_step 679, shared world
        # 706-996, tyres 997, car_dyn_pre 1261, car_step_rules 1402, sensor rays 1542)
        if line >= 1540: return "sensor rays"
        if line >= 1401: return "rules (lap timer, disable, reward, obs words)"
        if line >= 997: return "dynamics (forces, tyres, nearest segment)"
        if line >= 706: return "shared world (car-car, optional)"
        if line >= 679: return "body_step fast path"
        return "contact world (collide, solver, TOI driver, load/store)"
    if f == "ncg_b200.cu": return "kernel body (barriers, loads, stores)"
    return "other (" + str(f) + ")"
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": fn = row[1].split("(")[0][-60:]; continue
    if row[0] == "Line No": hdr = row; continue
    if row[0] == "": continue
    d = dict(zip(hdr, row))
    try:
        ie = int(d["Instructions Executed"]); te = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
    except Exception:
        continue
    for tab, key in ((fun, fn), (reg, region(cur, int(row[0])))):
        tab[key][0] += ie; tab[key][1] += te; tab[key][2] += smp
tot = sum(v[0] for v in reg.values()); tots = sum(v[2] for v in reg.values())
print(f"total warp-instructions {tot}  ({tot / steps / ctas:.0f} per CTA-step)   stall samples {tots}")
for name, tab in (("code region", reg), ("SASS function", fun)):
    print(f"--- by {name}")
    for k, v in sorted(tab.items(), key=lambda kv: -kv[1][0]):
        print(f"{k[:62]:62s} inst/CTA-step {v[0] / steps / ctas:8.0f} ({100 * v[0] / tot:4.1f} %)  lanes {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / max(tots, 1):4.1f} %")
