"""State -> renderer bridge: what ``CarEnv.render`` hands to the reference's ``Renderer.render_frame``
(/root/reference/src/car_env.py:1294-1401), rebuilt from the engine's per-car records -- so ``render_mode="human"`` can
draw a GPU-stepped env with the reference's own (pygame) renderer when that is importable.  Off the hot path; nothing
here touches the device.

``frame_kwargs`` is a pure function of host data (records + track table), which is how it is tested without pygame."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from . import constants as K
from . import info as I
from . import layout as L

R, F = L.R, L.F

# src/constants/car_specs.py:51-62
MULTI_CAR_COLORS = [(255, 0, 0), (0, 0, 255), (0, 255, 0), (255, 255, 0), (255, 0, 255), (0, 255, 255), (255, 128, 0),
                    (128, 0, 255), (255, 192, 203), (128, 128, 128)]


def track_progress(seg64: np.ndarray, x: float, y: float) -> float:
    """CarEnv._calculate_track_progress (car_env.py:1544-1611): float64, first strict minimum over the segments' chords."""
    best, bi, bcx, bcy = float("inf"), 0, 0.0, 0.0
    for i, s in enumerate(seg64):
        sx, sy, ex, ey = float(s[0]), float(s[1]), float(s[2]), float(s[3])
        dx, dy = ex - sx, ey - sy
        l2 = dx * dx + dy * dy
        if l2 < 1e-6:
            cx, cy = sx, sy
        else:
            t = max(0, min(1, ((x - sx) * dx + (y - sy) * dy) / l2))
            cx, cy = sx + t * dx, sy + t * dy
        d2 = (x - cx) ** 2 + (y - cy) ** 2
        if d2 < best:
            best, bi, bcx, bcy = d2, i, cx, cy
    total = 0.0
    for i in range(bi):
        s = seg64[i]
        total += math.sqrt((float(s[2]) - float(s[0])) ** 2 + (float(s[3]) - float(s[1])) ** 2)
    s = seg64[bi]
    return total + math.sqrt((bcx - float(s[0])) ** 2 + (bcy - float(s[1])) ** 2)


def race_positions(recs: np.ndarray, seg64: np.ndarray, track_length: float, car_names: Sequence[str]) -> list:
    """CarEnv._calculate_race_positions (car_env.py:1485-1542): (index, name, total progress, virtual laps, progress), leader first."""
    out = []
    for c in range(recs.shape[0]):
        fl = int(recs[c].view(np.uint32)[R["NCG_R_FLAGS"]])
        if fl & F["NCG_F_DISABLED"]:
            continue
        lt = I.lap_timing(recs[c])
        prog = track_progress(seg64, float(recs[c, R["NCG_R_X"]]), float(recs[c, R["NCG_R_Y"]]))
        laps = lt["lap_count"]
        if lt["is_timing"] and lt["has_crossed_startline"] and prog < track_length * 0.15 and lt["total_distance_traveled"] > track_length * 0.8:
            laps += 1
        out.append((c, car_names[c] if c < len(car_names) else f"Car {c}", laps * track_length + prog, laps, prog))
    out.sort(key=lambda x: (x[3], x[4]), reverse=True)
    return out


def best_lap_times(recs: np.ndarray, car_names: Sequence[str]) -> list:
    """CarEnv._get_best_lap_times_data (car_env.py:1613-1638)."""
    out = []
    for c in range(recs.shape[0]):
        fl = int(recs[c].view(np.uint32)[R["NCG_R_FLAGS"]])
        if fl & F["NCG_F_DISABLED"]:
            continue
        best = I.lap_timing(recs[c])["best_lap_time"]
        if best is not None:
            out.append((c, car_names[c] if c < len(car_names) else f"Car {c}", best))
    out.sort(key=lambda x: x[2])
    return out


def frame_kwargs(recs: np.ndarray, seg64: np.ndarray, track_length: float, car_names: Sequence[str], followed_car_index: int = 0,
                 last_rewards: Optional[Sequence[float]] = None, actions: Optional[np.ndarray] = None, reset_on_lap: bool = False,
                 track_file: Optional[str] = None, show_reward: bool = False) -> dict:
    """The keyword arguments of ``Renderer.render_frame`` as ``CarEnv.render`` builds them (car_env.py:1297-1401), from the
    (C, 128) records of one env.  ``debug_data`` and ``observation_info`` (the F1 / O overlays) are not bridged."""
    C = recs.shape[0]
    cars_data = [{"position": (float(recs[c, R["NCG_R_X"]]), float(recs[c, R["NCG_R_Y"]])), "angle": float(recs[c, R["NCG_R_ANGLE"]]),
                  "color": MULTI_CAR_COLORS[c] if c < len(MULTI_CAR_COLORS) else (255, 255, 255),
                  "name": car_names[c] if c < len(car_names) else f"Car {c}"} for c in range(C)]
    f = followed_car_index if followed_car_index < C else 0
    timing = I.lap_timing(recs[f])
    timing["car_name"] = car_names[f] if f < len(car_names) else f"Car {f}"
    step = int(recs[0].view(np.uint32)[R["NCG_R_STEP"]])
    reward_info = None
    if show_reward:
        reward_info = {"current_reward": float(last_rewards[f]) if last_rewards is not None else 0.0,
                       "cumulative_reward": float(recs[f, R["NCG_R_CUM_REWARD"]]), "show": True}
    return {
        "car_position": cars_data[f]["position"], "car_angle": cars_data[f]["angle"], "debug_data": None,
        "current_action": None if actions is None else actions[f], "lap_timing_info": timing, "reward_info": reward_info,
        "cars_data": cars_data, "followed_car_index": followed_car_index,
        "race_positions_data": race_positions(recs, seg64, track_length, car_names),
        "best_lap_times_data": best_lap_times(recs, car_names),
        "countdown_info": {"current_time": I.sim_time(step), "time_limit": K.TERMINATION_MAX_TIME if reset_on_lap else K.TRUNCATION_MAX_TIME,
                           "reset_on_lap": reset_on_lap},
        "observation_info": None, "track_file_name": track_file,
    }


def make_reference_renderer(track_file: str, render_fps: int = 60):
    """The reference's own Renderer over its own Track object.  Needs the reference tree on sys.path and pygame; raises
    ImportError otherwise (the engine itself never needs either)."""
    from src.constants import DEFAULT_WINDOW_SIZE          # noqa: the reference's modules, not this package's
    from src.renderer import Renderer
    from src.track_generator import TrackLoader
    return Renderer(window_size=DEFAULT_WINDOW_SIZE, render_fps=render_fps, track=TrackLoader().load_track(track_file))
