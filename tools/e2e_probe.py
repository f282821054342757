#!/usr/bin/env python
"""Where one host-buffer step goes: kernel alone, C-ABI pinned step, VectorEnv.step, and the bare copies."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.engine import Engine
from nascargymnasium_b200.vector_env import NascarVectorEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = 300
dev = torch.device("cuda:0")
eng = Engine(E, 1, tracks=["daytona"], auto_reset=True)
eng.reset_host()
acts = [torch.rand((E, 2), device=dev) * 2 - 1 for _ in range(64)]      # fresh actions every step: the random-policy regime
act = acts[0]
obs = torch.empty((E, 38), device=dev); rew = torch.empty(E, device=dev)
te = torch.empty(E, dtype=torch.uint8, device=dev); tr = torch.empty(E, dtype=torch.uint8, device=dev)
fin = torch.empty((E, 38), device=dev)
for _ in range(20): eng.step(act.view(-1), obs.view(-1), rew, te, tr, fin.view(-1))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n): eng.step(acts[i & 63].view(-1), obs.view(-1), rew, te, tr, fin.view(-1))
e1.record(); torch.cuda.synchronize()
print(f"E={E} kernel back-to-back (T=1, device buffers): {e0.elapsed_time(e1) / n * 1e3:.1f} us/step")
t0 = time.perf_counter()
for i in range(n):
    eng.step(acts[i & 63].view(-1), obs.view(-1), rew, te, tr, fin.view(-1)); torch.cuda.synchronize()
print(f"launch + sync each step (python): {(time.perf_counter() - t0) / n * 1e6:.1f} us/step")
eng.reset_host()
v = eng.pinned_views()
hacts = [np.random.uniform(-1, 1, v["actions"].shape).astype(np.float32) for _ in range(64)]
for i in range(20): v["actions"][...] = hacts[i]; eng.step_pinned(True)
t0 = time.perf_counter()
for i in range(n): v["actions"][...] = hacts[i & 63]; eng.step_pinned(True)
print(f"ncg_step_pinned (C ABI, H2D + kernel + D2H + sync): {(time.perf_counter() - t0) / n * 1e6:.1f} us/step")
venv = NascarVectorEnv(num_envs=E, track_file="tracks/daytona.track", num_cars=1)
venv.reset()
for i in range(20): venv.step(hacts[i])
t0 = time.perf_counter()
for i in range(n): venv.step(hacts[i & 63])
print(f"NascarVectorEnv.step (numpy in/out, mapped result buffers): {(time.perf_counter() - t0) / n * 1e6:.1f} us/step")
# the numpy statements of step() on their own
vv = venv.engine.pinned_views()
def tm(label, f):
    t0 = time.perf_counter()
    for _ in range(n): f()
    print(f"   {label}: {(time.perf_counter() - t0) / n * 1e6:.1f} us")
tm("actions -> pinned", lambda: vv["actions"].__setitem__(Ellipsis, np.asarray(hacts[0]).reshape(vv["actions"].shape)))
tm("obs.copy()", lambda: vv["obs"].reshape(venv._obs_shape).copy())
tm("rew.copy()", lambda: vv["reward"].reshape(venv._rew_shape).copy())
tm("flags astype(bool) x2", lambda: (vv["terminated"].astype(bool), vv["truncated"].astype(bool)))
tm("pinned_views()", lambda: venv.engine.pinned_views())
# bare copies
hp = torch.empty(E * 38 + E + 16, dtype=torch.float32).pin_memory(); dp = torch.empty_like(hp, device=dev)
ha = torch.empty(E * 2, dtype=torch.float32).pin_memory(); da = torch.empty_like(ha, device=dev)
s = torch.cuda.current_stream()
for _ in range(20): da.copy_(ha, non_blocking=True); hp.copy_(dp, non_blocking=True); s.synchronize()
t0 = time.perf_counter()
for _ in range(n): da.copy_(ha, non_blocking=True); hp.copy_(dp, non_blocking=True); s.synchronize()
print(f"bare H2D {ha.numel()*4} B + D2H {hp.numel()*4} B + sync: {(time.perf_counter() - t0) / n * 1e6:.1f} us")
t0 = time.perf_counter()
for _ in range(n): hp.copy_(dp, non_blocking=True); s.synchronize()
print(f"bare D2H + sync: {(time.perf_counter() - t0) / n * 1e6:.1f} us")
t0 = time.perf_counter()
for _ in range(n): s.synchronize()
print(f"bare sync: {(time.perf_counter() - t0) / n * 1e6:.1f} us")
