#!/usr/bin/env python
"""Needs a GPU and the NCG_TIMELINE variant build (NCG_VARIANT=tl NCG_DEFINES=-DNCG_TIMELINE).  Prints, for one CTA of a rollout
launch, how long each warp spends in each phase of a step (clock64 stamps; microseconds at the SM clock)."""
import ctypes, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200 import engine
from nascargymnasium_b200.engine import Engine

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
T = 100
eng = Engine(E, 1, tracks=["daytona"]); eng.reset_host()
o = torch.empty((T, E, 38), device="cuda")
for _ in range(30): eng.rollout(T, seed=0, mode=mode, obs_rollout=o.view(-1))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.rollout(T, seed=0, mode=mode, obs_rollout=o.view(-1)); b.record(); torch.cuda.synchronize()
print("launch %.1f us = %.2f us/step" % (a.elapsed_time(b) * 1e3, a.elapsed_time(b) * 1e3 / T))
tl = np.zeros((12, 128, 6), dtype=np.int64)
lib = engine.load_library()
lib.ncg_debug_timeline.argtypes = [ctypes.c_void_p]
assert lib.ncg_debug_timeline(tl.ctypes.data) == 0
mhz = 1965.0
S = slice(20, 90)
phys = tl[0, S]
d = lambda x: float(np.mean(x)) / mhz
print("physics warp: wait-empty %.2f  dynamics(->pose) %.2f  rules %.2f  env/finish %.2f  | step period %.2f us" % (
    d(phys[:, 1] - phys[:, 0]), d(phys[:, 2] - phys[:, 1]), d(phys[:, 3] - phys[:, 2]), d(phys[:, 4] - phys[:, 3]), d(np.diff(tl[0, 20:91, 0]))))
for w in range(1, 12):
    r = tl[w, S]
    if r[:, 0].min() == 0: continue
    print("ray warp %d: wait-pose %.2f  rays %.2f  wait-full %.2f  rows %.2f fence+arrive %.2f | pose lag (ray start - pose publish) %.2f" % (
        w, d(r[:, 1] - r[:, 0]), d(r[:, 2] - r[:, 1]), d(r[:, 3] - r[:, 2]), d(r[:, 5] - r[:, 3]), d(r[:, 4] - r[:, 5]), d(r[:, 1] - phys[:, 2])))

cc = np.zeros((4096, 4), dtype=np.int64)
lib.ncg_debug_cta_cycles.argtypes = [ctypes.c_void_p]
assert lib.ncg_debug_cta_cycles(cc.ctypes.data) == 0
n = int((cc[:, 0] > 0).sum())
k = cc[:n, 0] / mhz / T
print("per CTA (%d CTAs): us/step min %.2f  median %.2f  p90 %.2f  max %.2f   (CTA 5: %.2f)" % (n, k.min(), np.median(k), np.percentile(k, 90), k.max(), k[5] if n > 5 else 0))
pw, rw, rw2 = cc[:n, 1] / mhz / T, cc[:n, 2] / mhz / T, cc[:n, 3] / mhz / T
print("  waiting per step: physics warp (drained buffer) mean %.2f max %.2f | warp 1 mean %.2f min %.2f | warp 2 mean %.2f min %.2f" % (pw.mean(), pw.max(), rw.mean(), rw.min(), rw2.mean(), rw2.min()))
order = np.argsort(-k)[:8]
print("  slowest CTAs:", [(int(i), round(float(k[i]), 2), round(float(pw[i]), 2), round(float(rw[i]), 2)) for i in order])
