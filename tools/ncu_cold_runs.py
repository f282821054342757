#!/usr/bin/env python
"""Where does a warp jump over cold code?  From one `ncu --set full --import-source on` capture of ncg_step_kernel, list the
runs of SASS instructions inside the physics warp's step loop that (almost) never execute, with the source line of the
branch that skips them and the instruction-fetch stall samples on the instruction the jump lands on.  On B200 every
taken forward jump over >= ~30 instructions costs the warp ~100 cycles (stall_no_inst), whatever the distance.

    python tools/ncu_cold_runs.py gpurun_out/r01_full.ncu-rep [steps_per_launch=1000] [min_run=12]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
min_run = int(sys.argv[3]) if len(sys.argv) > 3 else 12


def ncu(page_args):
    return subprocess.run(f"ncu -i {rep} --page source --csv {page_args} 2>/dev/null", shell=True, capture_output=True, text=True).stdout


# address -> (file, line, text) from the CUDA-C correlated view
amap, cur, line, text = {}, None, None, None
for row in csv.reader(ncu("--print-source sass,cuda").splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        cur = row[1].split("/")[-1]
    elif row[0] in ("Function Name", "Line No"):
        continue
    elif row[0] != "":
        line, text = int(row[0]), row[1]
    elif len(row) > 3 and row[2].startswith("0x"):
        amap[row[2]] = (cur, line, text)

rows = list(csv.reader(ncu("--print-source sass").splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
src = lambda r: r[ix["Source"]].strip()
num = lambda r, k: int(r[ix[k]]) if r[ix[k]].isdigit() else 0

# the physics warp's loop: from the first predicated BAR.SYNC (buffer drained) to the second BAR.ARV (step complete)
bars = [n for n, r in enumerate(data) if "BAR.SYNC" in src(r) or "BAR.ARV" in src(r)]
lo = next(n for n in bars if src(data[n]).startswith("@"))
arv = [n for n in bars if "BAR.ARV" in src(data[n]) and n > lo]
hi = arv[1]
per_step = max(num(data[lo], "Instructions Executed") / steps, 1.0)        # = CTAs: executions per step of a once-per-step instruction
hot = [num(data[n], "Instructions Executed") >= 0.5 * per_step * steps for n in range(lo, hi)]
tot = collections.Counter()
for n in range(lo, hi):
    for k in ("# Samples", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_barrier", "stall_long_sb"):
        tot[k] += num(data[n], k)
    tot["inst"] += num(data[n], "Instructions Executed")
print(f"physics loop: {hi - lo} instruction slots, {sum(hot)} hot, {tot['inst'] / per_step / steps:.0f} executed per step; samples {tot['# Samples']}: "
      + ", ".join(f"{k[6:]} {100 * tot[k] / max(tot['# Samples'], 1):.1f}%" for k in tot if k.startswith("stall_")))
n = 0
while n < len(hot):
    if hot[n]:
        n += 1
        continue
    m = n
    while m < len(hot) and not hot[m]:
        m += 1
    if m - n >= min_run and m < len(hot):
        br, land, mid = data[lo + n - 1], data[lo + m], data[lo + (n + m) // 2]
        a, c = amap.get(br[ix["Address"]], ("?", 0, "")), amap.get(mid[ix["Address"]], ("?", 0, ""))
        print(f"{m - n:4d} cold | lands on no_inst {num(land, 'stall_no_inst'):4d} of {num(land, '# Samples'):4d} samples | skipped at {a[0]}:{a[1]} {a[2][:80]}")
        print(f"            cold body ~ {c[0]}:{c[1]} {c[2][:80]}")
    n = m
