#!/usr/bin/env python
"""Offline look at the cases tools/soak.py flagged (gpurun_out/soak_cases.npz): the same start record and action through
(a) the oracle, (b) the host compile of the device code (no FMA contraction).  If (b) agrees with the oracle and only the GPU
differs, the case is a rounding tie of a discrete decision (FMA contraction on the device), not a defect."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nascargymnasium_b200 import track as T, layout as L
from oracle import oracle as O
from tests import parity_util as P
R = L.R
d = np.load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "soak_cases.npz"))
names = list(T.BUILTIN_TRACK_NAMES)
n_tie = n_host_differs = 0
for i in range(len(d["track"])):
    name = names[int(d["track"][i])]
    rec, act = d["rec"][i], d["act"][i]
    pr = O.OracleEnv(T.builtin_track_text(name), num_cars=1)
    pr.set_state(P.record_to_oracle(rec))
    oo, ro, _, _ = pr.step(act[None])
    want = P.oracle_to_record(pr.get_state(), track_id=int(d["track"][i]))
    hc = P.HostCheckEnv(name)
    hc.records[0] = rec; hc.records.view(np.uint32)[0, R["NCG_R_TRACK"]] = 0
    a3 = np.array([[max(act[0], 0.0), max(-act[0], 0.0), act[1]]], dtype=np.float32)
    oh, rh, _, _, _ = hc.step(a3)
    touching = pr.num_contacts()[1] > 0
    b_host = P.compare_records(hc.records[0], want, touching=touching)
    b_gpu = P.compare_records(d["rec_after"][i], want, touching=touching)
    d_host = float(np.abs(oh[0] - oo[0]).max()); d_gpu = float(np.abs(d["obs"][i] - oo[0]).max())
    host_ok = not [x for x in b_host if x[0] != "NCG_R_TRACK"] and d_host < 1e-4
    n_tie += 1 if host_ok else 0; n_host_differs += 0 if host_ok else 1
    u = lambda r, k: int(np.asarray(r).view(np.uint32)[R[k]])
    print(f"case {i:2d} {name:14s} touching={int(touching)} speed={np.hypot(rec[R['NCG_R_VX']], rec[R['NCG_R_VY']]):6.2f} m/s  "
          f"contacts before/after(oracle, host, gpu) = {u(rec, 'NCG_R_NCONTACT') & 255}/({u(want, 'NCG_R_NCONTACT') & 255}, {u(hc.records[0], 'NCG_R_NCONTACT') & 255}, {u(d['rec_after'][i], 'NCG_R_NCONTACT') & 255})  "
          f"host compile vs oracle: {'SAME' if host_ok else [x[0] for x in b_host][:5]} (obs {d_host:.1e})   gpu vs oracle: {[x[0] for x in b_gpu][:5]} (obs {d_gpu:.1e})")
print(f"{n_tie} of {n_tie + n_host_differs} flagged cases: the host compile of the device code agrees with the oracle (the GPU's difference is a rounding tie); {n_host_differs}: the host compile differs too")
