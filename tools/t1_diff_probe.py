#!/usr/bin/env python
"""Needs a GPU.  ncg_step (device actions) vs ncg_rollout(T=1) on one engine in one steady state, 500 launches each."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.engine import Engine

out = []
for E in (4096, 65536):
    eng = Engine(E, 1, tracks=["daytona"])
    eng.reset_host()
    big = torch.empty((200, E, 38), device="cuda")
    for _ in range(15):
        eng.rollout(200, seed=0, obs_rollout=big.view(-1))
    obs = torch.empty((E, 38), device="cuda"); fin = torch.zeros((E, 38), device="cuda"); rew = torch.empty(E, device="cuda")
    te = torch.empty(E, dtype=torch.uint8, device="cuda"); tr = torch.empty(E, dtype=torch.uint8, device="cuda")
    acts = torch.rand((64, E, 2), device="cuda") * 2 - 1
    def timeit(fn, n=500):
        for i in range(20): fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for i in range(n): fn(i)
        b.record(); t_cpu = (time.perf_counter() - t0) * 1e6 / n
        torch.cuda.synchronize()
        return round(a.elapsed_time(b) * 1e3 / n, 2), round(t_cpu, 2)
    row = {"envs": E}
    row["rollout_T1_obs_last"] = timeit(lambda i: eng.rollout(1, seed=0, obs_last=obs.view(-1)))
    row["step_no_final"] = timeit(lambda i: eng.step(acts[i % 64].view(-1), obs.view(-1), rew, te, tr, None))
    row["step_final"] = timeit(lambda i: eng.step(acts[i % 64].view(-1), obs.view(-1), rew, te, tr, fin.view(-1)))
    row["step_same_action"] = timeit(lambda i: eng.step(acts[0].view(-1), obs.view(-1), rew, te, tr, None))
    row["rollout_T1_again"] = timeit(lambda i: eng.rollout(1, seed=0, obs_last=obs.view(-1)))
    out.append(row)
    eng.close()
print(json.dumps(out))
