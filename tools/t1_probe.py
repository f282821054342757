#!/usr/bin/env python
"""Needs a GPU.  Device time of single-step launches (T = 1, the path a learner drives through step_torch): `n` steps enqueued
back to back on one stream with device-resident actions, CUDA events around the whole run, no host synchronisation inside.

    python tools/t1_probe.py [--envs 4096 16384 65536] [--track daytona] [--discrete 0]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.vector_env import NascarVectorEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[4096, 16384, 65536])
    ap.add_argument("--track", default="daytona")
    ap.add_argument("--discrete", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1000)
    args = ap.parse_args()
    out = []
    for E in args.envs:
        v = NascarVectorEnv(E, track_file=f"tracks/{args.track}.track", discrete_action_space=bool(args.discrete))
        v.reset_torch()
        g = torch.Generator(device="cuda").manual_seed(0)
        # fresh actions every step (action_space.sample()): a short cycle of pre-drawn actions has a non-zero mean throttle and
        # steer per car, drives every car into a wall within a few hundred steps and then times the contact path instead
        n_sets = max(64, min(4096, (256 << 20) // (E * 8)))
        if args.discrete:
            acts = torch.randint(0, 5, (n_sets, E), device="cuda", generator=g, dtype=torch.int32)
        else:
            acts = torch.rand((n_sets, E, 2), device="cuda", generator=g) * 2 - 1
        for t in range(200):
            v.step_torch(acts[t % n_sets])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in range(args.steps):
            v.step_torch(acts[t % n_sets])
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / args.steps
        # where a single-step launch spends its time: the same loop without rays / without physics / without either
        # (NCG_DEBUG_SKIP is read at every launch; timing only, the results are not a simulation)
        skips = {}
        for sk in (1, 2, 3):
            os.environ["NCG_DEBUG_SKIP"] = str(sk)
            for t in range(50):
                v.step_torch(acts[t % n_sets])
            torch.cuda.synchronize()
            a.record()
            for t in range(500):
                v.step_torch(acts[t % n_sets])
            b.record()
            torch.cuda.synchronize()
            skips[{1: "no_rays", 2: "no_physics", 3: "neither"}[sk]] = a.elapsed_time(b) * 1e3 / 500
        os.environ.pop("NCG_DEBUG_SKIP", None)
        out.append({"envs": E, "track": args.track, "us_per_step": us, "car_steps_per_s": E / (us * 1e-6), "us_per_step_debug_skip": skips})
        v.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
