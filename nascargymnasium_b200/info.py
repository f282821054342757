"""Lazy ``info`` dict: the reference's key tree (CarEnv._get_multi_info, /root/reference/src/car_env.py:1160-1227)
materialised on the host from the engine's per-car records -- off the hot path, only when a caller asks."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import constants as K
from . import layout as L

R, F = L.R, L.F
TERMINATION_REASONS = (None, "all_cars_disabled", f"all_active_cars_low_reward (threshold: {K.TERMINATION_MIN_REWARD})", "time_limit",
                       "truncated")

# simulation_time after n env steps: the reference adds 1/60 in float64 once per step (car_env.py:573)
_TIMES = np.concatenate([[0.0], np.add.accumulate(np.full(K.TRUNCATION_STEPS + 8, 1.0 / 60.0, dtype=np.float64))])


def sim_time(steps: int) -> float:
    return float(_TIMES[steps]) if steps < len(_TIMES) else steps / 60.0


def format_time(t: Optional[float]) -> str:
    """LapTimer.format_time (lap_timer.py:320-345)."""
    if t is None or t < 0:
        return "--:--.---"
    total = int(t)
    return f"{total // 60:2d}:{total % 60:02d}.{int(round((t - total) * 1000)):03d}"


def _u(rec: np.ndarray, idx: int) -> int:
    return int(rec.view(np.uint32)[idx])


def lap_timing(rec: np.ndarray) -> dict:
    fl = _u(rec, R["NCG_R_FLAGS"])
    step, start = _u(rec, R["NCG_R_STEP"]), _u(rec, R["NCG_R_LAP_START"])
    timing = bool(fl & F["NCG_F_CROSSED"])
    cur = (sim_time(step - 1) - sim_time(start)) if timing and step > 0 and step - 1 >= start else 0.0
    last = float(rec[R["NCG_R_LAST_LAP"]]) if fl & F["NCG_F_HAS_LAST"] else None
    best = float(rec[R["NCG_R_BEST_LAP"]]) if fl & F["NCG_F_HAS_BEST"] else None
    return {
        "current_lap_time": cur, "last_lap_time": last, "best_lap_time": best, "lap_count": _u(rec, R["NCG_R_LAP_COUNT"]),
        "is_timing": timing, "has_crossed_startline": timing, "total_distance_traveled": float(rec[R["NCG_R_ODO"]]),
        "formatted_current": format_time(cur if timing else None), "formatted_last": format_time(last), "formatted_best": format_time(best),
    }


def performance(rec: np.ndarray, hist: Optional[np.ndarray] = None) -> dict:
    """Car.validate_performance (/root/reference/src/car.py:1060-1098) over Car.velocity_history: the (speed, dt) pairs of
    the last 600 update_physics calls of the episode.  `hist` is the engine's (600, 2) velocity ring of this car
    (Engine.velocity_history_host, kept with track_info=True); without it only the episode maximum is reported and the
    result is never valid."""
    out = {"max_speed_ms": K.CAR_MAX_SPEED_MS, "target_100kmh_ms": K.CAR_TARGET_100KMH_MS, "target_acceleration_time": K.CAR_ACCELERATION_0_100_KMH,
           "current_max_speed": 0.0, "estimated_0_100_time": 0.0, "performance_valid": False}
    if hist is None:
        out["current_max_speed"] = float(rec[R["NCG_R_MAX_SPEED"]])
        return out
    step = _u(rec, R["NCG_R_STEP"])                    # update_physics calls since Car.reset cleared the deque
    n = min(step, L.VEL_HISTORY)
    if n > K.PERFORMANCE_VALIDATION_MIN_SAMPLES:
        idx = np.arange(step - n, step) % L.VEL_HISTORY            # oldest sample of the window first
        v = np.asarray(hist, dtype=np.float32)[idx]
        speed = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(np.float64)   # b2Vec2.length: float32 arithmetic (car.py:928)
        mx = float(speed.max())
        out["current_max_speed"] = mx
        hit = np.flatnonzero(speed >= K.CAR_TARGET_100KMH_MS)
        if hit.size:
            out["estimated_0_100_time"] = float(np.add.accumulate(np.full(int(hit[0]) + 1, 1.0 / 60.0))[-1])
        t = out["estimated_0_100_time"]
        out["performance_valid"] = bool(mx >= K.CAR_MAX_SPEED_MS * K.PERFORMANCE_SPEED_TOLERANCE and
                                        (t <= K.CAR_ACCELERATION_0_100_KMH * K.PERFORMANCE_TIME_TOLERANCE if t > 0 else False))
    return out


def car_info(rec: np.ndarray, car_index: int, hist: Optional[np.ndarray] = None) -> dict:
    fl = _u(rec, R["NCG_R_FLAGS"])
    vx, vy = float(rec[R["NCG_R_VX"]]), float(rec[R["NCG_R_VY"]])
    speed = float(np.sqrt(np.float32(vx) * np.float32(vx) + np.float32(vy) * np.float32(vy)))
    return {
        "car_index": car_index, "disabled": bool(fl & F["NCG_F_DISABLED"]),
        "car_position": (float(rec[R["NCG_R_X"]]), float(rec[R["NCG_R_Y"]])), "car_speed_kmh": speed * 3.6, "car_speed_ms": speed,
        "on_track": bool(fl & F["NCG_F_ON_TRACK"]), "performance": performance(rec, hist), "lap_timing": lap_timing(rec),
        "cumulative_reward": float(rec[R["NCG_R_CUM_REWARD"]]), "cumulative_impact_force": float(rec[R["NCG_R_CUM_IMPACT"]]),
    }


def env_info(recs: np.ndarray, termination_reason: Optional[str] = None, followed_car_index: int = 0,
             hist: Optional[np.ndarray] = None) -> dict:
    """recs: (C,128) records of one env; hist: (C,600,2) velocity rings of its cars, or None."""
    C = recs.shape[0]
    step = _u(recs[0], R["NCG_R_STEP"])
    t = sim_time(step)
    h = (lambda c: hist[c]) if hist is not None else (lambda c: None)
    physics = [{"physics_steps": step, "simulation_time": t, "average_fps": 60.0, "bodies_in_world": 0, **performance(recs[c], h(c))}
               for c in range(C)]
    return {"simulation_time": t, "num_cars": C, "followed_car_index": followed_car_index, "termination_reason": termination_reason,
            "cars": [car_info(recs[c], c, h(c)) for c in range(C)], "physics": physics}
