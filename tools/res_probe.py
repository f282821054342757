#!/usr/bin/env python
"""NascarVectorEnv.step (numpy in / numpy out) with and without the resident step kernel: microseconds per step."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.vector_env import NascarVectorEnv

n = 2000
for E in [int(x) for x in (sys.argv[1:] or ["4096"])]:
    for res in ("0", "1"):
        os.environ["NCG_RESIDENT"] = res
        venv = NascarVectorEnv(num_envs=E, track_file="tracks/daytona.track", num_cars=1)
        venv.reset()
        rng = np.random.default_rng(0)
        acts = [rng.uniform(-1, 1, (E, 2)).astype(np.float32) for _ in range(n + 50)]      # fresh actions every step (a short cycle of them drives the cars into the walls)
        for i in range(50): venv.step(acts[i])
        l0 = venv.engine.launch_count
        t0 = time.perf_counter()
        for i in range(n): venv.step(acts[i + 50])
        dt = time.perf_counter() - t0
        print(f"E={E} NCG_RESIDENT={res}: step() {dt / n * 1e6:.1f} us  ({E * n / dt / 1e6:.1f} M car-steps/s), launches {venv.engine.launch_count - l0}, {venv.engine.resident_stats}", flush=True)
        venv.close()
