#!/usr/bin/env python
"""The reference's demo/random_demo.py loop (10 cars on nascar.track, reset_on_lap, random continuous actions) against the
CarEnv mirror -- the only change a caller makes is the import (and render_mode=None: the pygame window is outside the
accelerated path).

    python examples/random_demo.py [--steps 5000]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from nascargymnasium_b200.car_env import CarEnv          # reference: from src.car_env import CarEnv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--discrete", type=int, default=0)
    args = ap.parse_args()
    env = CarEnv(render_mode=None, track_file="tracks/nascar.track", reset_on_lap=True, num_cars=10,
                 discrete_action_space=bool(args.discrete))
    obs, info = env.reset()
    total_reward, episodes = 0.0, 0
    t0 = time.perf_counter()
    for step in range(args.steps):
        if env.check_quit_requested():
            break
        action = env.action_space.sample() if args.discrete else np.array(env.action_space.sample(), dtype=np.float32)
        obs, reward, terminated, truncated, info = env.step(action)
        total_reward += float(np.sum(reward))
        env.render()
        if terminated or truncated:
            print(f"   Episode terminated at step {step} ({info['termination_reason']}), total reward: {total_reward:.2f}")
            obs, info = env.reset()
            total_reward, episodes = 0.0, episodes + 1
    dt = time.perf_counter() - t0
    print(f"{args.steps} env steps x 10 cars in {dt:.2f} s = {args.steps * 10 / dt:.0f} car-steps/s through CarEnv.step "
          f"(one env per call; batch with NascarVectorEnv for throughput); {episodes} episodes")
    env.close()


if __name__ == "__main__":
    main()
