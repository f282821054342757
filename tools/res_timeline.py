#!/usr/bin/env python
"""Where a resident step goes on the device (variant build -DNCG_RES_TIMELINE): CTA 0's phases in clock cycles since it saw the
command.  NCG_VARIANT=restl NCG_DEFINES=-DNCG_RES_TIMELINE python tools/res_timeline.py [envs]"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NCG_RESIDENT"] = "1"
from nascargymnasium_b200 import engine
engine.build_library()
from nascargymnasium_b200.vector_env import NascarVectorEnv
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = 2000
venv = NascarVectorEnv(num_envs=E, track_file="tracks/daytona.track")
venv.reset()
warm = int(os.environ.get("WARM", "0"))
if warm:                                  # the random policy's dispersed steady state (device-side Philox actions, same distribution)
    import torch
    o = torch.empty((E, 38), device="cuda:0")
    venv.engine.rollout(warm, seed=1, mode=0, obs_last=o.view(-1)); torch.cuda.synchronize()
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (E, 2)).astype(np.float32) for _ in range(n + 50)]      # fresh actions every step (a short cycle of them drives the cars into the walls)
for i in range(n): venv.step(acts[i])
st = venv.engine.resident_stats
out = np.zeros(16, dtype=np.uint64)
lib = engine.load_library()
lib.ncg_debug_resident_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert lib.ncg_debug_resident_timeline(venv.engine._h, out.ctypes.data_as(ctypes.c_void_p)) == 0
names = ["actions arrived", "pose published", "rules done (before fence.sys)", "physics fence.sys done", "rays done", "FULL barrier passed",
         "rows stored", "ray warps met", "signaller fence.sys done", "dynamics done (pose in shared memory)", "fence.cta before the pose barrier done", "-", "-", "-", "-", "action load returned (before the slot pointers)"]
print(st)
scale = 1.965e3
if "TIMELINE2" in os.environ.get("NCG_DEFINES", ""):           # globaltimer (ns) since CTA 0 saw the command, ray warp stamps only
    names = ["command seen by this CTA (relay)", "pose barrier passed (ray warp)", "-", "-", "rays done", "FULL barrier passed", "rows stored", "-",
             "last CTA: its ray warps had met", "last CTA: GPU-scope fence done", "last CTA: counted (atomicAdd returned)"]
    scale = 1e3
for k, nm in enumerate(names):
    if nm == '-': continue
    print(f"  {nm:32s} {int(out[k]) / st['device_steps'] / scale:7.2f} us after the command was seen by CTA 0")
lanes = max(int(out[12]), 1)
print(f"  per lane: dynamics done {int(out[11]) / lanes / 1.965e3:.2f} us on average, {int(out[13]) / lanes * 100:.1f} % of the lanes later than 1.5 x lane 0, "
      f"{int(out[14]) / lanes * 100:.1f} % of the car-steps with broad-phase contacts")
