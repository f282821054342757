#!/usr/bin/env python
"""Soak of the resident step kernel: the same action stream through NascarVectorEnv.step with the resident kernel and with per-step
launches, every step's results hashed the moment step() returns (a result block that reached the host late or torn would change the
hash), hashes and final records compared.  python tools/res_soak.py [steps] [envs]"""
import os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.vector_env import NascarVectorEnv

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096


def run(resident):
    os.environ["NCG_RESIDENT"] = "1" if resident else "0"
    v = NascarVectorEnv(num_envs=E, track_file="tracks/daytona.track")
    v.reset()
    rng = np.random.default_rng(11)
    h = np.zeros(steps, dtype=np.uint32)
    dones = 0
    t0 = time.perf_counter()
    for t in range(steps):
        a = rng.uniform(-1, 1, (E, 2)).astype(np.float32)
        if (t // 2000) % 2:                              # every other 2000 steps the cars drive (wall contact, resets)
            a[:, 0] = np.abs(a[:, 0]) * 0.8 + 0.2
            a[:, 1] = a[:, 1] * 0.4 + 0.2
        o, r, te, tr, info = v.step(a)
        c = zlib.crc32(o); c = zlib.crc32(r, c); c = zlib.crc32(te, c); c = zlib.crc32(tr, c)
        if info:
            c = zlib.crc32(np.ascontiguousarray(info["final_obs_rows"]), c); dones += len(info["final_obs_index"])
        h[t] = c
    dt = time.perf_counter() - t0
    rec = v.engine.get_state_host().copy()
    st = v.engine.read_stats()
    res = v.engine.resident_stats
    launches = v.engine.launch_count
    v.close()
    return h, rec, st, res, launches, dt, dones


ha, ra, sa, resa, la, dta, da = run(True)
hb, rb, sb, resb, lb, dtb, db = run(False)
bad = np.flatnonzero(ha != hb)
print(f"{steps} steps x {E} envs: resident {dta:.1f} s ({resa['steps']} steps through the mailbox, {la} launches), launched {dtb:.1f} s ({lb} launches); "
      f"{da} finished episodes; contact steps {sa['contact_steps']}; steps whose results differ: {len(bad)}"
      f"{'' if not len(bad) else ' first at ' + str(bad[0])}; records equal: {np.array_equal(ra.view(np.uint32), rb.view(np.uint32))}; counters equal: {sa == sb}")
sys.exit(0 if not len(bad) and np.array_equal(ra.view(np.uint32), rb.view(np.uint32)) else 1)
