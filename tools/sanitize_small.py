"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): rollout + single steps + mapped step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.engine import Engine
from nascargymnasium_b200.vector_env import NascarVectorEnv
E = 70
eng = Engine(E, 1, tracks=["daytona", "martinsville"], auto_reset=True)
eng.reset_host(track_id=(np.arange(E) * 2 // E).astype(np.int32))
obs = torch.empty((6, E, 38), device="cuda:0")
eng.rollout(6, seed=1, mode=1, obs_rollout=obs.view(-1))
torch.cuda.synchronize()
for _ in range(3):
    eng.step_host(np.random.uniform(-1, 1, (E, 2)).astype(np.float32), want_final=True)
eng.close()
v = NascarVectorEnv(12, track_file="tracks/talladega.track", num_cars=10)
v.reset()
for _ in range(4):
    v.step(np.random.uniform(-1, 1, (12, 10, 2)).astype(np.float32))
v.close()
print("sanitize workload ok")
