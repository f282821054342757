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


def performance(rec: np.ndarray) -> dict:
    """Car.validate_performance keys (car.py:1060-1098).  The 600-sample velocity history is not kept on the device;
    ``current_max_speed`` is the episode maximum and the 0-100 km/h estimate is not available (0.0 => not valid)."""
    return {"max_speed_ms": K.CAR_MAX_SPEED_MS, "target_100kmh_ms": K.CAR_TARGET_100KMH_MS, "target_acceleration_time": K.CAR_ACCELERATION_0_100_KMH,
            "current_max_speed": float(rec[R["NCG_R_MAX_SPEED"]]), "estimated_0_100_time": 0.0, "performance_valid": False}


def car_info(rec: np.ndarray, car_index: int) -> dict:
    fl = _u(rec, R["NCG_R_FLAGS"])
    vx, vy = float(rec[R["NCG_R_VX"]]), float(rec[R["NCG_R_VY"]])
    speed = float(np.sqrt(np.float32(vx) * np.float32(vx) + np.float32(vy) * np.float32(vy)))
    return {
        "car_index": car_index, "disabled": bool(fl & F["NCG_F_DISABLED"]),
        "car_position": (float(rec[R["NCG_R_X"]]), float(rec[R["NCG_R_Y"]])), "car_speed_kmh": speed * 3.6, "car_speed_ms": speed,
        "on_track": bool(fl & F["NCG_F_ON_TRACK"]), "performance": performance(rec), "lap_timing": lap_timing(rec),
        "cumulative_reward": float(rec[R["NCG_R_CUM_REWARD"]]), "cumulative_impact_force": float(rec[R["NCG_R_CUM_IMPACT"]]),
    }


def env_info(recs: np.ndarray, termination_reason: Optional[str] = None, followed_car_index: int = 0) -> dict:
    """recs: (C,128) records of one env."""
    C = recs.shape[0]
    step = _u(recs[0], R["NCG_R_STEP"])
    t = sim_time(step)
    physics = [{"physics_steps": step, "simulation_time": t, "average_fps": 60.0, "bodies_in_world": 0, **performance(recs[c])}
               for c in range(C)]
    return {"simulation_time": t, "num_cars": C, "followed_car_index": followed_car_index, "termination_reason": termination_reason,
            "cars": [car_info(recs[c], c) for c in range(C)], "physics": physics}
