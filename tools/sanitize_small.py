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
for i in range(8):                 # (through the resident step kernel; the state read in the middle ends and restarts it)
    v.step(np.random.uniform(-1, 1, (12, 10, 2)).astype(np.float32))
    if i == 3:
        v.engine.set_state_host(v.engine.get_state_host())
v.close()
# interleaved env -> track map (slot list), random-track mode with a redraw, start pose, Box2D 2.3.0 contact form
w = NascarVectorEnv(48, track_file=None, discrete_action_space=True)
w.reset(seed=3)
w.engine.set_state_host(w.engine.get_state_host())
rec = w.engine.get_state_host()
from nascargymnasium_b200 import layout as L
rec.view(np.uint32)[:, L.R["NCG_R_STUCK_STEPS"]] = 599                      # every env finishes on the next step and moves to another track
w.engine.set_state_host(rec)
for _ in range(3):
    w.step(np.zeros(48, dtype=np.int64))
w.close()
e2 = Engine(40, 1, tracks=["nascar"], contacts=2, start_position=(30.0, -4.0), start_angle=0.35, track_info=True)
e2.reset_host()
o2 = torch.empty((40, 38), device="cuda:0")
e2.rollout(30, seed=2, mode=1, obs_last=o2.view(-1))
torch.cuda.synchronize()
e2.close()
print("sanitize workload ok")
