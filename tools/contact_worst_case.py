#!/usr/bin/env python
"""TEST INFRASTRUCTURE (needs a GPU): the measured worst case, per record field, of engine vs oracle on teacher-forced
steps that solve wall contacts -- the evidence behind the contact-step tolerances of tests/parity_util.py.

    python tools/contact_worst_case.py [--cases 1500] > gpurun_out/contact_worst_case.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nascargymnasium_b200 import layout as L  # noqa: E402
from tests import parity_util as P  # noqa: E402

R = L.R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=1500)
    args = ap.parse_args()
    from nascargymnasium_b200.engine import Engine
    fields = {name: cnt for name, (cnt, _, _) in P.FLOAT_FIELDS.items()}
    worst = {name: {"abs": 0.0, "rel": 0.0, "scale": 0.0} for name in list(fields) + ["manifold_impulses", "obs[0:22]", "obs[22:38]", "reward"]}
    n_touch = n_all = flags_bad = ids_bad = 0
    for track, kind, seed in (("martinsville", "drive", 1), ("michigan", "drive", 4), ("nascar", "full", 0), ("trioval", "drive", 6),
                              ("daytona", "drive", 8)):
        recs, act3, raws, exp = P.collect_cases(track, args.cases, kind=kind, seed=seed, every=3)
        m = len(recs)
        eng = Engine(m, 1, tracks=[track], auto_reset=False)
        eng.reset_host()
        eng.set_state_host(recs)
        obs, rew, te, tr, _ = eng.step_host(np.array(raws, dtype=np.float32))
        got = eng.get_state_host()
        eng.close()
        for i in range(m):
            n_all += 1
            if exp["touching"][i] <= 0:
                continue
            n_touch += 1
            w = exp["records"][i]
            # discrete outcome first: contact list, touching set, point counts, feature ids.  A case where they differ is a
            # threshold tie (a separation within rounding of b2_polygonRadius or of the reference-face tolerance); it is
            # counted and left out of the per-field figures, which describe steps that solved the same constraints.
            ncw = P._u(w, R["NCG_R_NCONTACT"])
            same = P._u(got[i], R["NCG_R_NCONTACT"]) == ncw and P._u(got[i], R["NCG_R_MANIFOLD_PC"]) == P._u(w, R["NCG_R_MANIFOLD_PC"])
            nt = bin((ncw >> 16) & 0xFFF).count("1")
            for k in range(nt if same else 0):
                M = R["NCG_R_MANIFOLD"] + 6 * k
                same = same and P._u(got[i], M) == P._u(w, M) and P._u(got[i], M + 1) == P._u(w, M + 1)
            if not same:
                ids_bad += 1
                continue
            for name, cnt in fields.items():
                a, b = got[i, R[name]:R[name] + cnt].astype(np.float64), w[R[name]:R[name] + cnt].astype(np.float64)
                d = np.abs(a - b)
                k = int(d.argmax())
                if d[k] > worst[name]["abs"]:
                    worst[name].update(abs=float(d[k]), scale=float(abs(b[k])))
                big = np.abs(b) > 1e-2
                if big.any():
                    worst[name]["rel"] = max(worst[name]["rel"], float((d[big] / np.abs(b[big])).max()))
            for k in range(nt):
                M = R["NCG_R_MANIFOLD"] + 6 * k
                a, b = got[i, M + 2:M + 6].astype(np.float64), w[M + 2:M + 6].astype(np.float64)
                d = np.abs(a - b)
                j = int(d.argmax())
                if d[j] > worst["manifold_impulses"]["abs"]:
                    worst["manifold_impulses"].update(abs=float(d[j]), scale=float(abs(b[j])))
                big = np.abs(b) > 10.0
                if big.any():
                    worst["manifold_impulses"]["rel"] = max(worst["manifold_impulses"]["rel"], float((d[big] / np.abs(b[big])).max()))
            for key, sl in (("obs[0:22]", slice(0, 22)), ("obs[22:38]", slice(22, 38))):
                d = float(np.abs(obs[i, sl] - exp["obs"][i, sl]).max())
                worst[key]["abs"] = max(worst[key]["abs"], d)
            worst["reward"]["abs"] = max(worst["reward"]["abs"], abs(float(rew[i]) - float(exp["reward"][i])))
            if bool(te[i]) != bool(exp["terminated"][i]) or bool(tr[i]) != bool(exp["truncated"][i]):
                flags_bad += 1
    out = {"what": "engine (CUDA, float32, FMA contraction on) vs oracle (gcc, -ffp-contract=off) on teacher-forced single steps that "
                   "solve wall contacts; worst absolute difference per record field, the magnitude of the value it occurred on, and "
                   "the worst relative difference", "cases": n_all, "touching_cases": n_touch, "contact_list_or_feature_id_mismatches": ids_bad,
           "flag_mismatches": flags_bad, "worst": worst}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
