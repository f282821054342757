"""The state -> renderer bridge (render_bridge.frame_kwargs): the keyword arguments the reference's CarEnv.render hands to
Renderer.render_frame (/root/reference/src/car_env.py:1294-1401, 1485-1542, 1613-1638), rebuilt from engine records.  Pure host
code: the records come from the reference-pinned oracle replay of golden trajectories."""
import os

import numpy as np
import pytest

from nascargymnasium_b200 import render_bridge as RB
from nascargymnasium_b200 import track as T
from oracle import oracle as O
from tests import parity_util as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _replay(name, upto):
    with np.load(os.path.join(GOLD, f"traj_{name}.npz")) as z:
        g = {k: z[k] for k in z.files}
    C = int(g["num_cars"])
    env = O.OracleEnv(T.builtin_track_text(str(g["track"])), num_cars=C, reset_on_lap=bool(g["reset_on_lap"]), discrete=bool(g["discrete"]))
    env.reset()
    for t in range(upto):
        env.step(g["actions"][t])
        if g["did_reset"][t]:
            env.reset(fresh=False)
    recs = np.stack([P.oracle_to_record(env.get_state(c)) for c in range(C)])
    return g, env, recs


def test_frame_kwargs_after_a_completed_lap():
    g, env, recs = _replay("martinsville_laps", 1870)          # the lap completes at step 1877 and the env resets: stop before
    tab = T.get_track_table("martinsville")
    kw = RB.frame_kwargs(recs, tab.seg64, tab.track.total_length, ["Car 0"], 0, last_rewards=[0.25], actions=np.array([[0.5, 0.1]]),
                         reset_on_lap=True, track_file="tracks/martinsville.track", show_reward=True)
    assert set(kw) == {"car_position", "car_angle", "debug_data", "current_action", "lap_timing_info", "reward_info", "cars_data",
                       "followed_car_index", "race_positions_data", "best_lap_times_data", "countdown_info", "observation_info",
                       "track_file_name"}
    s = env.get_state(0)
    assert kw["car_position"] == pytest.approx((float(np.float32(s[P.S["S_X"]])), float(np.float32(s[P.S["S_Y"]]))))
    assert kw["cars_data"][0]["color"] == (255, 0, 0) and kw["cars_data"][0]["name"] == "Car 0"
    assert kw["lap_timing_info"]["car_name"] == "Car 0" and kw["lap_timing_info"]["is_timing"]
    assert kw["countdown_info"] == {"current_time": pytest.approx(env.sim_time, abs=1e-12), "time_limit": 60.0, "reset_on_lap": True}
    assert kw["reward_info"]["current_reward"] == 0.25 and kw["reward_info"]["show"] is True
    (idx, name, total, laps, prog), = kw["race_positions_data"]
    want = O.lib().orc_env_progress(env._h, float(recs[0, P.R["NCG_R_X"]]), float(recs[0, P.R["NCG_R_Y"]]))
    assert prog == pytest.approx(want, abs=1e-9)                # the reference-pinned float64 chord search
    # just before the line with > 80 % of the lap driven: the reference counts a virtual lap for the standings
    assert laps == (1 if (prog < 0.15 * tab.track.total_length and kw["lap_timing_info"]["total_distance_traveled"] > 0.8 * tab.track.total_length) else 0)
    assert total == pytest.approx(laps * tab.track.total_length + prog)
    assert kw["best_lap_times_data"] == []                      # no lap completed yet


def test_standings_of_a_three_car_env_are_sorted_like_the_reference_sorts_them():
    g, env, recs = _replay("trioval_3cars_laplimit", 1500)
    tab = T.get_track_table("trioval")
    names = ["A", "B", "C"]
    kw = RB.frame_kwargs(recs, tab.seg64, tab.track.total_length, names, followed_car_index=2)
    rows = kw["race_positions_data"]
    disabled = {c for c in range(3) if int(recs[c].view(np.uint32)[P.R["NCG_R_FLAGS"]]) & P.F["NCG_F_DISABLED"]}
    assert {r[0] for r in rows} == set(range(3)) - disabled
    assert rows == sorted(rows, key=lambda x: (x[3], x[4]), reverse=True)
    for r in rows:
        assert r[4] == pytest.approx(O.lib().orc_env_progress(env._h, float(recs[r[0], P.R["NCG_R_X"]]), float(recs[r[0], P.R["NCG_R_Y"]])), abs=1e-9)
        assert r[1] == names[r[0]]
    assert kw["followed_car_index"] == 2 and kw["lap_timing_info"]["car_name"] == "C" and kw["reward_info"] is None
    assert kw["car_angle"] == pytest.approx(float(recs[2, P.R["NCG_R_ANGLE"]]))


def test_human_render_mode_asks_for_the_reference_renderer():
    from nascargymnasium_b200.car_env import CarEnv
    env = CarEnv(render_mode="human", track_file="tracks/nascar.track")          # constructing needs neither a GPU nor pygame
    with pytest.raises(NotImplementedError, match="pygame Renderer"):
        env.render()
    assert CarEnv(track_file="tracks/nascar.track").render() is None
