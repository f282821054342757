#!/usr/bin/env python
"""Needs a GPU.  Long free runs of the engine (all 8 tracks, both synthetic policies, single- and ten-car envs) with periodic
spot checks against the oracle AT THE STATES THE FREE RUN REACHES: every `every` launches the records of a random sample of cars
are handed to the oracle (teacher forcing), both take one more step with the same action, and records / observation / reward /
flags are compared with the parity tolerances of tests/parity_util.py.  Also checked all along: finite state, observation
ranges, no capacity overflow.  Prints one JSON summary (kept as profiles/rNN_soak.json).

    python tools/soak.py [--minutes 3] [--envs 4096] [--sample 48]
"""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nascargymnasium_b200.engine import Engine
from nascargymnasium_b200 import track as T, layout as L
from oracle import oracle as O
from tests import parity_util as P

R = L.R
ap = argparse.ArgumentParser()
ap.add_argument("--minutes", type=float, default=3.0)
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--sample", type=int, default=48)
ap.add_argument("--steps-per-launch", type=int, default=250)
args = ap.parse_args()
names = list(T.BUILTIN_TRACK_NAMES)
texts = [T.builtin_track_text(n) for n in names]
rng = np.random.default_rng(0)
out = {"runs": []}
dump = []
budget = args.minutes * 60.0 / 4
for C, mode in ((1, 1), (1, 0), (10, 1), (10, 0)):
    E = args.envs if C == 1 else max(args.envs // 10, 64)
    eng = Engine(E, C, tracks=names, auto_reset=True)
    tid = (np.arange(E) * len(names) // E).astype(np.int32)
    eng.reset_host(track_id=tid)
    last = torch.empty((E * C, 38), device="cuda:0")
    probes = [O.OracleEnv(texts[k], num_cars=1) for k in range(len(names))]
    t0, launches, checked, touching_checked, worst = time.time(), 0, 0, 0, {}
    fails = []
    while time.time() - t0 < budget:
        eng.rollout(args.steps_per_launch, seed=11 + launches, mode=mode, obs_last=last.view(-1))
        launches += 1
        torch.cuda.synchronize()
        o = last.cpu().numpy()
        assert np.isfinite(o).all() and o.min() >= -1.0 and o.max() <= 1.0 and (o[:, 22:] >= 0).all()
        if launches % 4:
            continue
        # ---- spot check: one teacher-forced step from the reached states, engine (single-step launch) vs oracle
        recs = eng.get_state_host().copy()
        cars = rng.choice(E * C, size=min(args.sample, E * C), replace=False)
        acts = rng.uniform(-1, 1, size=(E * C, 2)).astype(np.float32)
        if mode == 1:
            acts[:, 0] = rng.uniform(0.2, 1.0, size=E * C); acts[:, 1] = rng.uniform(-0.2, 0.6, size=E * C)
        obs_g, rew_g, te_g, tr_g, _ = eng.step_host(acts.reshape(E, C, 2) if C > 1 else acts)
        recs_after = eng.get_state_host()
        obs_g = obs_g.reshape(E * C, 38)
        for gc in cars:
            e = gc // C
            if te_g[e] or tr_g[e]:
                continue                                   # the env finished in this step: its record is already the reset state
            rec = recs[gc]
            pr = probes[tid[e]]
            pr.set_state(P.record_to_oracle(rec))
            oo, ro, _, _ = pr.step(acts[gc:gc + 1])
            try:
                want = P.oracle_to_record(pr.get_state(), track_id=int(tid[e]))
            except OverflowError:
                continue                                   # the oracle state exceeds a record cap (counted by the engine as overflow)
            touching = pr.num_contacts()[1] > 0
            b = P.compare_records(recs_after[gc], want, touching=touching)
            d22 = np.abs(obs_g[gc, :22] - oo[0, :22])
            if touching:
                d22[6] = max(0.0, d22[6] - 1e-4)           # obs[6] = omega / 10: see compare_records
            d_obs, d_ray = float(d22.max()), float(np.abs(obs_g[gc, 22:] - oo[0, 22:]).max())
            d_rew = abs(float(rew_g.reshape(-1)[gc]) - float(ro[0]))
            worst["obs"] = max(worst.get("obs", 0.0), d_obs); worst["rays"] = max(worst.get("rays", 0.0), d_ray)
            if b or d_obs > 1e-4 or d_ray > 1e-3 or d_rew > 1e-4 + 1e-4 * abs(float(ro[0])):
                fails.append({"launch": launches, "car": int(gc), "track": names[tid[e]], "touching": bool(touching), "fields": [x[0] for x in b][:6],
                              "d_obs": d_obs, "d_rays": d_ray, "d_reward": d_rew})
                dump.append((int(tid[e]), rec.copy(), acts[gc].copy(), recs_after[gc].copy(), obs_g[gc].copy()))
            checked += 1; touching_checked += 1 if touching else 0
    st = eng.read_stats()
    out["runs"].append({"cars_per_env": C, "envs": E, "policy": "driving" if mode == 1 else "random", "launches": launches,
                        "car_steps": int(st["car_steps"]), "contact_steps": int(st["contact_steps"]), "toi_events": int(st["toi_events"]),
                        "episodes": int(st["episodes"]), "laps": int(st["laps"]), "overflow": int(st["overflow"]),
                        "spot_checks": checked, "spot_checks_in_contact": touching_checked, "spot_check_failures": fails[:10], "n_failures": len(fails),
                        "worst_obs_diff": worst.get("obs"), "worst_ray_diff": worst.get("rays"), "seconds": round(time.time() - t0, 1)})
    eng.close()
if dump:      # the failing cases, for the offline look (tools/soak_cases.py): start record, action, the engine's record and observation after the step
    np.savez(os.path.join(ROOT, "gpurun_out", "soak_cases.npz"), track=np.array([d[0] for d in dump]), rec=np.array([d[1] for d in dump]),
             act=np.array([d[2] for d in dump]), rec_after=np.array([d[3] for d in dump]), obs=np.array([d[4] for d in dump]))
print(json.dumps(out))
