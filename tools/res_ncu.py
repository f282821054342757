#!/usr/bin/env python
"""Workload for an ncu capture of the resident step kernel: one 3000-step rollout launch (steady state), then four steps through
NascarVectorEnv.step.  Under ncu a launch returns when the kernel has ended, so each of the first three steps is one resident launch
(the step waiting in its mailbox, then the idle time: NCG_RESIDENT_IDLE_US=30 keeps the spin short).
ncu --set full -k regex:ncg_step_kernel --launch-skip 2 --launch-count 1 ... python tools/res_ncu.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCG_RESIDENT_IDLE_US", "30")
from nascargymnasium_b200.vector_env import NascarVectorEnv
E = 4096
v = NascarVectorEnv(num_envs=E, track_file="tracks/daytona.track")
v.reset()
o = torch.empty((E, 38), device="cuda:0")
v.engine.rollout(3000, seed=1, mode=0, obs_last=o.view(-1)); torch.cuda.synchronize()
rng = np.random.default_rng(0)
for _ in range(4):
    v.step(rng.uniform(-1, 1, (E, 2)).astype(np.float32))
print("launches", v.engine.launch_count, v.engine.resident_stats)
v.close()
