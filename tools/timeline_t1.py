#!/usr/bin/env python
"""Needs a GPU and the NCG_TIMELINE variant build.  Where a SINGLE-STEP launch (T = 1, what step_torch issues) spends its time
inside the kernel, for one CTA: prologue (table staging, record load) / dynamics / rays / rows / epilogue (record store)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200 import engine
from nascargymnasium_b200.engine import Engine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = Engine(E, 1, tracks=["daytona"]); eng.reset_host()
o = torch.empty((1, E, 38), device="cuda")
for _ in range(3000): eng.rollout(1, seed=0, obs_rollout=o.view(-1))
torch.cuda.synchronize()
lib = engine.load_library()
lib.ncg_debug_launch_ns.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.ncg_debug_launch_ns(None, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200): eng.rollout(1, seed=0, obs_rollout=o.view(-1))
b.record(); torch.cuda.synchronize()
print("E=%d: %.2f us per single-step launch (back to back)" % (E, a.elapsed_time(b) * 1e3 / 200))
lib = engine.load_library()
tl = np.zeros((12, 128, 6), dtype=np.int64); cc = np.zeros((4096, 5), dtype=np.int64)
lib.ncg_debug_timeline.argtypes = [ctypes.c_void_p]; lib.ncg_debug_cta_cycles.argtypes = [ctypes.c_void_p]
assert lib.ncg_debug_timeline(tl.ctypes.data) == 0 and lib.ncg_debug_cta_cycles(cc.ctypes.data) == 0
mhz = 1965.0
k0, tot = cc[5, 4], cc[5, 0]
ph = tl[0, 0]; end_sync = tl[11, 127, 0]
us = lambda c: c / mhz
print("CTA 5: kernel %.2f us | prologue (start -> physics loop) %.2f | dynamics -> pose %.2f | rules %.2f | env/finish %.2f" % (
    us(tot), us(ph[0] - k0), us(ph[2] - ph[1]), us(ph[3] - ph[2]), us(ph[4] - ph[3])))
for w in (1, 2, 3):
    r = tl[w, 0]
    print("  ray warp %d: loop start %.2f after kernel start | pose seen at %.2f | rays %.2f | full seen at %.2f | rows done at %.2f" % (
        w, us(r[0] - k0), us(r[1] - k0), us(r[2] - r[1]), us(r[3] - k0), us(r[5] - k0)))
print("  all warps past the final __syncthreads at %.2f; record store + exit %.2f" % (us(end_sync - k0), us(k0 + tot - end_sync)))
n = int((cc[:, 0] > 0).sum()); k = cc[:n, 0] / mhz
print("per CTA kernel time: min %.2f median %.2f max %.2f us over %d CTAs" % (k.min(), np.median(k), k.max(), n))

ln = np.zeros((4096, 2), dtype=np.uint64)
lib.ncg_debug_launch_ns(ln.ctypes.data, 0)
ok = ln[:, 1] > 0
st, en = ln[ok, 0].astype(np.int64), ln[ok, 1].astype(np.int64)
order = np.argsort(st); st, en = st[order], en[order]
dur = (en - st) / 1e3; gap = (st[1:] - en[:-1]) / 1e3
print("globaltimer over %d launches: first CTA start -> last CTA end median %.2f us; gap to the next launch's first CTA median %.2f us (p10 %.2f, p90 %.2f)" % (
    len(st), np.median(dur), np.median(gap), np.percentile(gap, 10), np.percentile(gap, 90)))
