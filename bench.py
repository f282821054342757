#!/usr/bin/env python
"""bench.py -- car-steps/sec of the batched CarEnv stepping path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference on all host cores

A "step" is one CarEnv.step() of every env of the workload (one batched car-step with full 16-ray
observations, reward, termination and same-step auto-reset).  Workload at any N: BASELINE.json configs[1]
per GPU -- 4096 single-car envs on daytona.track, continuous actions, random policy (weak scaling).

`value`   : whole-job car-steps/s, state and actions resident in HBM (rollout kernel, Philox actions on device,
            observations of every step written to HBM), CUDA-event time on the launching stream, max over ranks.
`e2e`     : the same metric through the CarEnv-facing host API (VectorEnv.step with numpy actions): host->device
            copy of the actions and device->host copy of obs/reward/flags inside the timed region, every step.
`roofline`: algorithmic bytes (832 B per car-step) / kernel duration vs measured HBM peak.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_CAR_STEP = 832     # 83 state words read + written, 8 B action, 152 B obs, 4 B reward, 4 B flags
METRIC = "car-steps/sec (16-ray sensors) at 4096-65536 envs, 1/2/4/8 B200 vs host CPU"
UNIT = "car-steps/s"


def profiled_traffic(steps_per_launch: int, n_cars: int):
    """DRAM bytes per launch of the rollout kernel from the committed `ncu --set full` capture (profiles/), or None when
    the capture was taken at another launch shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)
        if int(t["steps_per_launch"]) == steps_per_launch and int(t["cars"]) == n_cars:
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ----------------------------------------------------------------------------- Philox4x32-10 (same stream as the device)
def philox_uniform2(seed: int, cars: np.ndarray, step: int):
    """u0,u1 in [0,1) for (car, step): mirrors ncg::action_synthetic (csrc/ncg_car.cuh)."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c0 = cars.astype(np.uint64); c1 = np.full_like(c0, step); c2 = np.zeros_like(c0); c3 = np.zeros_like(c0)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & mask
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    to01 = lambda x: (x >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return to01(c0), to01(c1)


def synthetic_actions(seed: int, cars: np.ndarray, step: int) -> np.ndarray:
    u0, u1 = philox_uniform2(seed, cars, step)
    return np.stack([np.float32(2.0) * u0 - np.float32(1.0), np.float32(2.0) * u1 - np.float32(1.0)], axis=1).astype(np.float32)


# ----------------------------------------------------------------------------- CPU arm (oracle = port of the reference)
def _cpu_worker(args):
    track, car_base, n_steps, seed, warm = args
    from oracle import oracle as O
    from nascargymnasium_b200 import track as T
    env = O.OracleEnv(T.builtin_track_text(track))
    env.reset()
    cars = np.array([car_base], dtype=np.int64)
    acts = [synthetic_actions(seed, cars, s) for s in range(warm + n_steps)]
    for s in range(warm):
        _, _, te, tr = env.step(acts[s])
        if te or tr:
            env.reset(fresh=False)
    t0 = time.perf_counter()
    for s in range(warm, warm + n_steps):
        _, _, te, tr = env.step(acts[s])
        if te or tr:
            env.reset(fresh=False)
    return n_steps, time.perf_counter() - t0


def cpu_throughput(track: str, steps_per_proc: int, procs: int, seed: int = 0, warm: int = 600):
    """car-steps/s of the oracle with `procs` independent single-car envs, one per process."""
    jobs = [(track, i, steps_per_proc, seed, warm) for i in range(procs)]
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 5 ms; nvidia-smi as a fallback)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons, self._stop, self._t = index, [], [], set(), threading.Event(), None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for bit, nm in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")):
            if r & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        s = [x.strip() for x in out.split(",")]
        if s[0].replace(".", "").isdigit():
            self.sm.append(float(s[0]))
        if len(s) > 1 and s[1].replace(".", "").isdigit():
            self.mx.append(float(s[1]))
        for k, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if len(s) > 3 + k and s[3 + k].lower().startswith("active"):
                self.reasons.add(nm)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.005 if self._nvml else 0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self._nvml else "nvidia-smi"}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from nascargymnasium_b200.engine import Engine
    from nascargymnasium_b200.vector_env import NascarVectorEnv

    E, C, track = args.envs, args.cars, args.track
    N = E * C
    K, Wm = args.steps, args.warmup
    # steps per launch: long rollouts amortise the cold start of a launch (after the L2 flush the kernel's own code and the
    # track tables come from HBM: ~0.16 ms per launch, measured); capped so the rollout buffer stays below 2 GB
    T = max(1, min(args.steps_per_launch, K, int(2e9 // (N * 38 * 4))))
    from nascargymnasium_b200 import track as TR
    tracks = list(TR.BUILTIN_TRACK_NAMES) if track == "all" else [track]
    track_id = (np.arange(E, dtype=np.int64) * len(tracks) // E).astype(np.int32)      # contiguous blocks of envs per track
    eng = Engine(E, C, tracks=tracks, discrete=False, auto_reset=True, device=local)
    eng.reset_host(track_id=track_id)
    obs_roll = torch.empty((T, N, 38), dtype=torch.float32, device=dev)
    rew_roll = torch.empty((T, N), dtype=torch.float32, device=dev)
    done_roll = torch.empty((T, E), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launch(n):
        eng.rollout(n, seed=args.seed, mode=args.mode, obs_rollout=obs_roll[:n].reshape(-1) if n != T else obs_roll.reshape(-1),
                    reward_rollout=rew_roll[:n].reshape(-1) if n != T else rew_roll.reshape(-1),
                    done_rollout=done_roll[:n].reshape(-1) if n != T else done_roll.reshape(-1))

    # warm-up: W steps (at least 3 launches)
    done = 0
    for _ in range(max(3, (Wm + T - 1) // T)):
        launch(T)
        done += T
    barrier()
    eng.read_stats(reset=True)
    launches0 = eng.launch_count
    chunks = [T] * (K // T) + ([K % T] if K % T else [])
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in chunks]
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for (e0, e1), n in zip(ev, chunks):
            flush.zero_()                         # L2 flush between timed launches (not inside the event pair)
            e0.record()
            launch(n)
            e1.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    gpu_launches = eng.launch_count - launches0
    stats = eng.read_stats(reset=True)
    assert stats["car_steps"] == N * K, (stats, N, K)
    t = torch.tensor([kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kern_ms_max = float(t.item())
    value = world * N * K / (kern_ms_max / 1e3)

    # ---- e2e: host buffers through the CarEnv-facing API, H2D + D2H every step
    venv = NascarVectorEnv(num_envs=E, track_file=None if track == "all" else f"tracks/{track}.track", num_cars=C, device=local)
    venv.reset()
    Ke = args.e2e_steps
    cars = np.arange(N, dtype=np.int64) + rank * N
    acts = [synthetic_actions(args.seed, cars, s).reshape(E, C, 2) if C > 1 else synthetic_actions(args.seed, cars, s) for s in range(Ke + 5)]
    for s in range(5):
        venv.step(acts[s])
    barrier()
    l0 = venv.engine.launch_count
    t0 = time.perf_counter()
    for s in range(5, Ke + 5):
        o, r, te, tr, info = venv.step(acts[s])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    gpu_launches += venv.engine.launch_count - l0
    te2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te2, op=dist.ReduceOp.MAX)
    e2e_value = world * N * Ke / float(te2.item())
    h2d = N * 2 * 4
    d2h = N * 38 * 4 + N * 4 + 2 * E

    peak, peak_src = measured_peak()
    per_launch_s = (kern_ms / 1e3) / len(chunks)
    achieved = ALGO_BYTES_PER_CAR_STEP * N * (K / len(chunks)) / per_launch_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": kern_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{E} batched {'single-car' if C == 1 else str(C) + '-car'} envs on {track}.track per GPU, continuous "
                               f"actions {'~ U[-1,1]^2' if args.mode == 0 else 'driving distribution tb~U[0.2,1], steer~U[-0.2,0.6]'} (Philox on "
                               f"device), 16-ray observations written every step, same-step auto-reset",
                   "envs_per_gpu": E, "cars_per_env": C, "track": track, "steps_per_launch": T,
                   "l2": "flushed (256 MiB memset) between timed launches; the working set itself is L2-resident by construction",
                   "rays_per_lane": os.environ.get("NCG_RAYS_PER_LANE", "auto")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                "api": "NascarVectorEnv.step(numpy actions) -> numpy obs/reward/terminated/truncated; actions are copied into "
                       "page-locked memory the kernel reads across PCIe, results are written by the kernel into page-locked host "
                       "buffers and returned without a further copy"},
        "gpu_launches": int(gpu_launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": profiled_traffic(T, N) if args.mode == 0 and track == "daytona" else None,
                     "peak_source": peak_src, "kernel": "ncg_step_kernel", "algorithmic_bytes_per_car_step": ALGO_BYTES_PER_CAR_STEP,
                     "car_steps_per_launch": N * (K / len(chunks)), "avg_launch_ms": per_launch_s * 1e3},
        "clocks": clk.summary(),
        "counters": {k: int(v) if k != "return_sum" else float(v) for k, v in stats.items()},
        "wall_s_timed_region": t_wall,
    }
    if rank == 0 and world == 1 and args.sweep:
        # the same kernel at larger batches (informational: the metric is quoted at 4096-65536 envs); short runs
        line["batch_sweep"] = []
        eng.close()
        for Es, trk in ((8192, track), (16384, track), (65536, track), (65536, "all")):
            names = list(TR.BUILTIN_TRACK_NAMES) if trk == "all" else [trk]
            e2 = Engine(Es, C, tracks=names, discrete=False, auto_reset=True, device=local)
            e2.reset_host(track_id=(np.arange(Es, dtype=np.int64) * len(names) // Es).astype(np.int32))
            o2 = torch.empty((50, Es * C, 38), dtype=torch.float32, device=dev)
            for _ in range(6):
                e2.rollout(50, seed=args.seed, mode=args.mode, obs_rollout=o2.reshape(-1))
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                e2.rollout(50, seed=args.seed, mode=args.mode, obs_rollout=o2.reshape(-1))
            b.record()
            torch.cuda.synchronize()
            line["batch_sweep"].append({"envs": Es, "track": trk, "value": Es * C * 500 / (a.elapsed_time(b) / 1e3), "unit": UNIT,
                                        "steps": 500, "note": "rollout kernel, 50 steps per launch, no L2 flush"})
            e2.close()
            del o2
    if rank == 0 and world == 1:
        t0 = time.perf_counter()
        v = cpu_throughput(tracks[0], args.cpu_steps, 1, seed=args.seed)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"1 single-car env on {track}.track, {args.cpu_steps} steps after 600 warm-up, same Philox action "
                                          f"stream, oracle/ncg_oracle.cpp single thread ({time.perf_counter() - t0:.1f} s)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU implementation of the path: real box2d-py/gymnasium are not installable offline, so this
    arm times the oracle port (oracle/ncg_oracle.cpp) with one process per host core."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    K, Wm = args.steps, args.warmup
    # each "step" is a bounded sample: every core advances one single-car env by cpu_steps/ K ... keep total ~20-40 s
    per_proc = max(200, min(args.cpu_steps, 40000))
    t0 = time.perf_counter()
    v = cpu_throughput("daytona" if args.track == "all" else args.track, per_proc, cores, seed=args.seed)
    dt = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": 1e3 * args.envs * args.cars / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64/f32", "data": "synthetic",
        "config": {"workload": f"{args.envs} batched single-car envs on {args.track}.track per GPU, continuous actions ~ U[-1,1]^2",
                   "envs_per_gpu": args.envs, "cars_per_env": args.cars, "track": args.track},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x 1 single-car env x {per_proc} steps after 600 warm-up ({dt:.1f} s wall); "
                                   "box2d-py/gymnasium absent offline, so the oracle port stands in for the reference CarEnv"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30000)
    ap.add_argument("--warmup", type=int, default=3000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--cars", type=int, default=1)
    ap.add_argument("--track", default="daytona", help="a built-in track name, or 'all' = the 8 .track files in equal blocks of envs")
    ap.add_argument("--steps-per-launch", type=int, default=1000)
    ap.add_argument("--e2e-steps", type=int, default=2000)
    ap.add_argument("--cpu-steps", type=int, default=40000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--sweep", type=int, default=1, help="1: also time the rollout kernel at 8192/16384/65536 envs (N=1 only)")
    ap.add_argument("--mode", type=int, default=0, help="synthetic action distribution: 0 = action_space.sample() (the metric), "
                    "1 = 'driving' (tb~U[0.2,1], steer~U[-0.2,0.6]: laps, wall contacts, episodes)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
