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
    ap.add_argument("--steps", type=int, default=2000)
    args = ap.parse_args()
    out = []
    for E in args.envs:
        v = NascarVectorEnv(E, track_file=f"tracks/{args.track}.track", discrete_action_space=bool(args.discrete))
        v.reset_torch()
        g = torch.Generator(device="cuda").manual_seed(0)
        if args.discrete:
            acts = torch.randint(0, 5, (64, E), device="cuda", generator=g, dtype=torch.int32)
        else:
            acts = torch.rand((64, E, 2), device="cuda", generator=g) * 2 - 1
        for t in range(200):
            v.step_torch(acts[t % 64])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in range(args.steps):
            v.step_torch(acts[t % 64])
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / args.steps
        # the same step captured in a CUDA graph: the launch as the GPU sees it, without the Python / ctypes call
        static_a = acts[0].clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                v.step_torch(static_a)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            v.step_torch(static_a)
        for _ in range(50):
            g.replay()
        torch.cuda.synchronize()
        a.record()
        for t in range(args.steps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        us_g = a.elapsed_time(b) * 1e3 / args.steps
        out.append({"envs": E, "track": args.track, "us_per_step": us, "us_per_step_graph": us_g, "car_steps_per_s": E / (us * 1e-6),
                    "car_steps_per_s_graph": E / (us_g * 1e-6)})
        v.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
