#!/usr/bin/env python
"""TEST INFRASTRUCTURE: how sensitive is this world's contact behaviour to WHICH Box2D 2.3.x box2d-py 2.3.8 bundles?

The only place where the 2.3.x releases differ for a box-vs-static-box world is b2CollidePolygons (oracle/b2lite.h,
g_collide_variant): 2.3.0 finds the reference edge by a hill climb over b2EdgeSeparation and flips the reference face with
k_relativeTol 0.98 / k_absoluteTol 0.001; 2.3.1+ uses a brute-force max and k_tol = 0.1 * b2_linearSlop.  This script

  1. draws random car poses around the walls of every track and counts how often the two manifolds differ, and
  2. replays the golden action streams (tests/golden/traj_*.npz, recorded from the reference's own CarEnv) and 20 000 steps
     of the contact-heavy "driving" distribution per track under both variants, free-running, and reports the first step at
     which the observations differ at all and the number of contact steps before it.

    python tools/b2_version_study.py [--poses 2000000] > profiles/r02_b2_version_study.json
"""
import argparse
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nascargymnasium_b200 import track as T  # noqa: E402
from oracle import oracle as O  # noqa: E402


def replay(track, actions, variant, num_cars=1, reset_on_lap=False, discrete=False, did_reset=None, start=(0.0, 0.0, 0.0)):
    O.set_b2_variant(variant)
    env = O.OracleEnv(T.builtin_track_text(track), num_cars=num_cars, reset_on_lap=reset_on_lap, discrete=discrete,
                      start_position=(start[0], start[1]), start_angle=start[2])
    env.reset()
    obs, touching = [], []
    for t, a in enumerate(actions):
        o, r, te, tr = env.step(a if not discrete else a.astype(np.int64))
        obs.append(o.copy())
        touching.append(sum(env.num_contacts(c)[1] for c in range(num_cars)))
        if (did_reset is not None and did_reset[t]) or (did_reset is None and (te or tr)):
            env.reset(fresh=False)
    O.set_b2_variant(0)
    return np.array(obs), np.array(touching)


def compare(track, actions, **kw):
    o0, t0 = replay(track, actions, 0, **kw)
    o1, t1 = replay(track, actions, 1, **kw)
    diff = np.flatnonzero((o0 != o1).reshape(len(o0), -1).any(axis=1))
    first = int(diff[0]) if len(diff) else None
    upto = first if first is not None else len(o0)
    return {"steps": len(o0), "contact_steps": int((t0 > 0).sum()), "first_differing_step": first,
            "contact_steps_before_it": int((t0[:upto] > 0).sum()),
            "max_abs_obs_difference": float(np.abs(o0 - o1).max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=2_000_000)
    ap.add_argument("--drive-steps", type=int, default=20000)
    args = ap.parse_args()
    out = {"what": "Box2D 2.3.0 vs 2.3.1 b2CollidePolygons on this world (oracle/b2lite.h g_collide_variant)", "poses": {}, "goldens": {}, "driving": {}}
    tot = None
    for name in T.BUILTIN_TRACK_NAMES:
        env = O.OracleEnv(T.builtin_track_text(name))
        c = O.b2_variant_study(env, args.poses // len(T.BUILTIN_TRACK_NAMES), seed=11)
        out["poses"][name] = c
        tot = c if tot is None else {k: (max(tot[k], v) if k.startswith("worst") else tot[k] + v) for k, v in c.items()}
    touching = max(1, tot["touching"])
    tot["manifold_differs_per_million_touching"] = 1e6 * (tot["touching_differs"] + tot["reference_face_differs"] + tot["point_count_differs"]
                                                          + tot["feature_ids_differ"]) / touching
    tot["bits_differ_per_million_touching"] = 1e6 * tot["same_features_bits_differ"] / touching
    out["poses"]["all_tracks"] = tot
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "traj_*.npz"))):
        with np.load(path) as z:
            g = {k: z[k] for k in z.files}
        sp = g["start_pose"]
        out["goldens"][os.path.basename(path)[5:-4]] = compare(
            str(g["track"]), g["actions"], num_cars=int(g["num_cars"]), reset_on_lap=bool(g["reset_on_lap"]), discrete=bool(g["discrete"]),
            did_reset=g["did_reset"], start=(float(sp[0]), float(sp[1]), float(sp[2])))
    rng = np.random.default_rng(3)
    for name in T.BUILTIN_TRACK_NAMES:
        acts = np.stack([rng.uniform(0.2, 1.0, size=args.drive_steps), rng.uniform(-0.2, 0.6, size=args.drive_steps)], axis=1).astype(np.float32)
        out["driving"][name] = compare(name, acts[:, None, :])
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
