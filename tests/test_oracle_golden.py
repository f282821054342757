"""Pins the oracle (oracle/ncg_oracle.cpp) against fixtures produced by the REFERENCE's own code.

tests/golden/traj_*.npz were written by oracle/gen_golden.py: the reference's unmodified src/car_env.py stepping over
the stand-in Box2D of oracle/refshim (backed by oracle/b2lite.h).  Replaying the recorded actions through the oracle's
native step must reproduce the reference's observations, rewards and flags: that pins every line of the oracle's
restatement of car.py / tyre*.py / lap_timer.py / car_env.py / distance_sensor.py / track_generator.py.
(The rigid-body step underneath both is b2lite: Box2D itself stays unpinned, see oracle/b2lite.h.)"""
import glob
import json
import os

import numpy as np
import pytest

from nascargymnasium_b200 import constants as K
from nascargymnasium_b200 import track as T
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAJ = sorted(glob.glob(os.path.join(GOLD, "traj_*.npz")))
# tests/golden_real/: the same trajectories recorded over REAL box2d-py (`python oracle/gen_golden.py --real-box2d` on a machine
# that has it; none exists offline).  When present they are replayed too, contacts included: that pins oracle/b2lite.h itself.
# NCG_BOX2D=2.3.0 replays them under the 2.3.0 form of b2CollidePolygons.
TRAJ += sorted(glob.glob(os.path.join(os.path.dirname(GOLD), "golden_real", "traj_*.npz")))
if os.environ.get("NCG_BOX2D", "").strip() == "2.3.0":
    O.set_b2_variant(1)
LAYOUT = O.state_layout()


def test_fixtures_present():
    assert len(TRAJ) >= 8
    for f in ("constants.json", "tracks.json", "unit_kats.json"):
        assert os.path.exists(os.path.join(GOLD, f))


@pytest.mark.parametrize("path", TRAJ, ids=[("real_" if "golden_real" in p else "") + os.path.basename(p)[5:-4] for p in TRAJ])
def test_oracle_reproduces_reference_trajectory(path):
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}            # NpzFile re-inflates on every access
    track, C = str(g["track"]), int(g["num_cars"])
    sp = g["start_pose"]
    env = O.OracleEnv(T.builtin_track_text(track), num_cars=C, reset_on_lap=bool(g["reset_on_lap"]), discrete=bool(g["discrete"]),
                      start_position=(float(sp[0]), float(sp[1])), start_angle=float(sp[2]))
    obs0 = env.reset()
    assert np.abs(obs0 - g["obs0"]).max() < 1e-7
    n = len(g["actions"])
    worst_obs = worst_rew = 0.0
    for t in range(n):
        a = g["actions"][t]
        obs, rew, te, tr = env.step(a if not g["discrete"] else a.astype(np.int64))
        assert te == bool(g["terminated"][t]) and tr == bool(g["truncated"][t]), f"flags differ at step {t}"
        worst_obs = max(worst_obs, float(np.abs(obs - g["obs"][t]).max()))
        worst_rew = max(worst_rew, float(np.abs(rew - g["reward"][t]).max()))
        assert worst_obs < 2e-6 and worst_rew < 1e-5, (t, worst_obs, worst_rew)
        assert abs(env.sim_time - float(g["sim_time"][t])) < 1e-12
        assert (env.termination_reason or "") == str(g["reason"][t])
        if t % 20 == 0 or te or tr or t == n - 1:
            for c in range(C):
                s = env.get_state(c)
                assert int(s[LAYOUT["S_LAPCOUNT"]]) == int(g["lap_count"][t][c])
                assert bool(s[LAYOUT["S_DISABLED"]]) == bool(g["disabled"][t][c])
                assert env.on_track(c) == bool(g["on_track"][t][c])
        if g["did_reset"][t]:
            env.reset(fresh=False)
    print(os.path.basename(path), "steps", n, "max |dobs|", worst_obs, "max |drew|", worst_rew)


def test_unit_kats_from_reference_classes():
    with open(os.path.join(GOLD, "unit_kats.json")) as f:
        k = json.load(f)
    for case in k["tyres"]:
        dt, along, alat, speed, slip = case["args"]
        # the reference object keeps state across the 5 updates; the oracle entry point is single-shot, so check step 1
        out = O.kat_tyres(case["friction"], dt, along, alat, speed, slip)
        want = np.array(case["out"][0])
        assert np.allclose(out, want, rtol=1e-13, atol=1e-13)
    rpm = 1000.0
    for thr, want in k["rpm"]:
        rpm = O.kat_rpm(rpm, thr)
        assert abs(rpm - want) < 1e-9
    assert k["torque"][0][2] == pytest.approx(574.0)


def test_constants_match_reference():
    with open(os.path.join(GOLD, "constants.json")) as f:
        ref = json.load(f)
    pairs = {
        "CAR_MASS": K.CAR_MASS, "CAR_LENGTH": K.CAR_LENGTH, "CAR_WIDTH": K.CAR_WIDTH, "CAR_WHEELBASE": K.CAR_WHEELBASE,
        "CAR_MAX_TORQUE": K.CAR_MAX_TORQUE, "CAR_MAX_POWER": K.CAR_MAX_POWER, "CAR_MAX_SPEED_MS": K.CAR_MAX_SPEED_MS,
        "CAR_MOMENT_OF_INERTIA": K.CAR_MOMENT_OF_INERTIA, "DRAG_CONSTANT": K.DRAG_CONSTANT, "CAR_FRICTION": K.CAR_FRICTION,
        "CAR_RESTITUTION": K.CAR_RESTITUTION, "BOX2D_WALL_FRICTION": K.WALL_FRICTION, "BOX2D_WALL_RESTITUTION": K.WALL_RESTITUTION,
        "BOX2D_TIME_STEP": K.TIME_STEP, "BOX2D_VELOCITY_ITERATIONS": K.VELOCITY_ITERATIONS, "BOX2D_POSITION_ITERATIONS": K.POSITION_ITERATIONS,
        "ENGINE_IDLE_RPM": K.ENGINE_IDLE_RPM, "ENGINE_REDLINE_RPM_RANGE": K.ENGINE_RPM_RANGE, "ENGINE_RPM_RESPONSE_RATE": K.ENGINE_RPM_RESPONSE_RATE,
        "FINAL_DRIVE_RATIO": K.FINAL_DRIVE_RATIO, "WHEEL_RADIUS": K.WHEEL_RADIUS, "MAX_STEERING_ANGLE": K.MAX_STEERING_ANGLE_DEG,
        "ROLLING_RESISTANCE_FORCE": K.ROLLING_RESISTANCE_FORCE, "STATIC_LOAD_PER_TYRE": K.STATIC_LOAD_PER_TYRE, "MAX_TYRE_LOAD": K.MAX_TYRE_LOAD,
        "MAX_BRAKE_DECELERATION_G": K.MAX_BRAKE_DECEL, "STEERING_TORQUE_MULTIPLIER": K.STEER_TORQUE_MULT, "STEERING_ANGULAR_DAMPING": K.ANGULAR_DAMPING,
        "MAX_LATERAL_FORCE": K.MAX_LATERAL_FORCE, "VELOCITY_ALIGNMENT_FORCE_FACTOR": K.ALIGN_FORCE_FACTOR,
        "TYRE_START_TEMPERATURE": K.TYRE_START_TEMP, "TYRE_IDEAL_TEMPERATURE_MIN": K.TYRE_IDEAL_MIN, "TYRE_IDEAL_TEMPERATURE_MAX": K.TYRE_IDEAL_MAX,
        "COLLISION_FORCE_THRESHOLD": K.COLLISION_FORCE_THRESHOLD, "STUCK_SPEED_THRESHOLD": K.STUCK_SPEED, "STUCK_TIME_THRESHOLD": K.STUCK_TIME,
        "STUCK_DISTANCE_THRESHOLD": K.STUCK_DISTANCE, "STUCK_EXTENDED_TIME_THRESHOLD": K.STUCK_EXTENDED_TIME,
        "BACKWARD_MOVEMENT_THRESHOLD": K.BACKWARD_PENALTY_START, "BACKWARD_DISABLE_THRESHOLD": K.BACKWARD_DISABLE,
        "INSTANT_DISABLE_IMPACT_THRESHOLD": K.INSTANT_DISABLE_IMPACT, "CUMULATIVE_DISABLE_IMPACT_THRESHOLD": K.CUMULATIVE_DISABLE_IMPACT,
        "SENSOR_NUM_DIRECTIONS": K.NUM_SENSORS, "SENSOR_MAX_DISTANCE": K.SENSOR_MAX_DISTANCE, "NORM_MAX_POSITION": K.NORM_POS,
        "NORM_MAX_VELOCITY": K.NORM_VEL, "NORM_MAX_ANGULAR_VEL": K.NORM_ANGVEL, "NORM_MAX_TYRE_TEMP": K.NORM_TYRE_TEMP,
        "NORM_MAX_TYRE_WEAR": K.NORM_TYRE_WEAR, "NORM_MAX_TYRE_LOAD": K.NORM_TYRE_LOAD,
        "REWARD_DISTANCE_MULTIPLIER": K.REWARD_DISTANCE, "PENALTY_PER_STEP": K.PENALTY_PER_STEP, "PENALTY_BACKWARD_PER_METER": K.PENALTY_BACKWARD_PER_M,
        "PENALTY_WALL_COLLISION_PER_STEP": K.PENALTY_WALL, "PENALTY_DISABLED": K.PENALTY_DISABLED, "TERMINATION_MIN_REWARD": K.TERMINATION_MIN_REWARD,
        "TERMINATION_MAX_TIME": K.TERMINATION_MAX_TIME, "TRUNCATION_MAX_TIME": K.TRUNCATION_MAX_TIME, "TRACK_WALL_THICKNESS": K.WALL_THICKNESS,
        "MAX_CARS": K.MAX_CARS,
    }
    for name, mine in pairs.items():
        assert name in ref, name
        assert ref[name] == pytest.approx(mine, rel=1e-15), name
