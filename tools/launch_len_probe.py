#!/usr/bin/env python
"""Needs a GPU.  Per-step device time of the rollout kernel as a function of the steps per launch T (same engine, same
steady state, no L2 flush): what a launch costs on top of its steps."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.engine import Engine

out = []
for E in (4096, 65536):
    eng = Engine(E, 1, tracks=["daytona"])
    eng.reset_host()
    obs = torch.empty((200, E, 38), device="cuda")
    for _ in range(15):
        eng.rollout(200, seed=0, obs_rollout=obs.view(-1))
    torch.cuda.synchronize()
    row = {"envs": E}
    for T in (1, 2, 4, 8, 20, 50, 200):
        n = max(3, 2000 // T)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            eng.rollout(T, seed=0, obs_rollout=obs[:T].reshape(-1))
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            eng.rollout(T, seed=0, obs_rollout=obs[:T].reshape(-1))
        b.record()
        torch.cuda.synchronize()
        row[f"T={T}"] = round(a.elapsed_time(b) * 1e3 / (n * T), 2)
    out.append(row)
    eng.close()
print(json.dumps(out))
