#!/usr/bin/env python
"""Assemble profiles/<round>_* from a gpurun_out/<tag>_* capture set (tools/gpu_profile.sh).
   python tools/make_profiles.py r01b r01"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def run(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout

def raw(rep):
    rows = list(csv.reader(run(f"ncu -i {rep} --page raw --csv 2>/dev/null").splitlines()))
    return rows[0], rows[1], rows[2]

out = []
for name, rep, what in (("rollout kernel (T = 1000 steps per launch, steady state)", f"{G}/{tag}_full.ncu-rep", "full"),
                        ("single-step kernel (T = 1, results written to mapped host memory: the e2e path)", f"{G}/{tag}_full_t1.ncu-rep", "full_t1"),
                        ("rollout kernel at 65536 envs (T = 100 steps per launch; ray queue, launch shape picked by launch_step)", f"{G}/{tag}_full_big.ncu-rep", "full_big")):
    if not os.path.exists(rep):
        continue
    open(f"/tmp/{what}_raw.csv", "w").write(run(f"ncu -i {rep} --page raw --csv 2>/dev/null"))
    open(f"/tmp/{what}_src.csv", "w").write(run(f"ncu -i {rep} --page source --csv --print-source sass,cuda 2>/dev/null"))
    out.append(f"## {name}\n\n```\n" + run(f"python {ROOT}/tools/ncu_summary.py /tmp/{what}_raw.csv") + "```\n")
    out.append("Top source lines by executed warp instructions (ncu source page, -lineinfo):\n\n```\n" +
               run(f"python {ROOT}/tools/ncu_lines.py /tmp/{what}_src.csv 40") + "```\n")
    if what == "full":
        h, u, r = raw(rep)
        g = lambda k: float(r[h.index(k)]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[u[h.index(k)]]
        bench = json.load(open(f"{G}/{tag}_plain_short.json"))
        json.dump({"source": f"profiles/{rnd}_ncu_summary.md ({tag}_full.ncu-rep, ncu --set full, one rollout launch)",
                   "steps_per_launch": bench["config"]["steps_per_launch"], "cars": bench["config"]["envs_per_gpu"] * bench["config"]["cars_per_env"],
                   "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum")},
                  open(f"{P}/roofline_traffic.json", "w"), indent=1)
# launch list
rows = [r for r in csv.reader(open(f"{G}/{tag}_launches.csv")) if len(r) > 10 and r[0].isdigit()]
with open(f"{P}/{rnd}_launches.csv", "w") as f:
    f.write("id,kernel,block,grid,duration_ns\n")
    for r in rows:
        f.write(f"{r[0]},\"{r[4].split('(')[0]}\",\"{r[7]}\",\"{r[8]}\",{r[-1]}\n")
tot = sum(float(r[-1]) for r in rows)
ours = sum(float(r[-1]) for r in rows if "ncg_" in r[4])
roll = [float(r[-1]) for r in rows if "ncg_step_kernel" in r[4] and float(r[-1]) > 3e5]
single = [float(r[-1]) for r in rows if "ncg_step_kernel" in r[4] and float(r[-1]) <= 3e5]
hdr = f"""# ncu evidence, round {rnd[1:]} (capture set gpurun_out/{tag}_*, B200, clocks unlocked: --clock-control none)

Command profiled: `python bench.py --steps 1000 --warmup 1000 --e2e-steps 20 --cpu-steps 200 --sweep 0` (same code path as the default
bench line, shorter).  Launch list: `{rnd}_launches.csv` ({len(rows)} launches; ncu serialises launches and runs them cold).
Share of GPU time in the list: ncg_* kernels {100 * ours / tot:.1f} % (the rest is torch's 256 MiB L2-flush memset between timed
launches, outside the CUDA-event pairs).  Rollout launches (1000 steps): {len(roll)} x {sum(roll) / max(len(roll), 1) / 1e3:.0f} us
= {sum(roll) / max(len(roll), 1) / 1000 / 1e3:.2f} us per step; single-step launches of the e2e path: {len(single)} x
{sum(single) / max(len(single), 1) / 1e3:.1f} us (these write 647 KB per launch into mapped host memory).

"""
open(f"{P}/{rnd}_ncu_summary.md", "w").write(hdr + "\n".join(out))
for f in (f"{tag}_bench.json", f"{tag}_bench_ref.json", f"{tag}_gpu.txt", f"{tag}_pytest.log"):
    if os.path.exists(f"{G}/{f}"):
        open(f"{P}/{rnd}_{f.split('_', 1)[1]}", "w").write(open(f"{G}/{f}").read())
print(open(f"{P}/{rnd}_ncu_summary.md").read()[:1500])
