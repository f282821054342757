#!/usr/bin/env python
"""Needs a GPU.  Per-step device time of the rollout kernel as a function of the steps since reset (launches of 50 steps):
how much the cost of a step depends on where the cars are."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.engine import Engine

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
eng = Engine(E, 1, tracks=["daytona"])
eng.reset_host()
obs = torch.empty((50, E, 38), device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(81)]
ev[0].record()
for i in range(80):
    eng.rollout(50, seed=0, mode=mode, obs_rollout=obs.view(-1))
    ev[i + 1].record()
torch.cuda.synchronize()
print(json.dumps({"envs": E, "mode": mode, "us_per_step_by_launch_of_50": [round(ev[i].elapsed_time(ev[i + 1]) * 1e3 / 50, 2) for i in range(80)],
                  "stats": eng.read_stats()}))
