#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generate tests/golden/* by running the REFERENCE's own code in the build container.

    python oracle/gen_golden.py                # needs /root/reference (read-only); writes tests/golden/
    python oracle/gen_golden.py --real-box2d   # on a machine with `pip install box2d-py==2.3.8`: the same trajectories
                                               # over REAL Box2D, written to tests/golden_real/ (pins the Box2D half)

What runs: the reference's unmodified modules (src/car_env.py, car.py, car_physics.py, tyre*.py, lap_timer.py,
track_generator.py, distance_sensor.py, constants/*) imported from /root/reference, over the stand-in modules in
oracle/refshim/ (pygame: no-op; gymnasium: Env + spaces; Box2D: pybox2d API subset backed by oracle/b2lite.h).
So every number the reference computes in Python is the reference's; the rigid-body step underneath is this
repo's Box2D restatement (real box2d-py is not installable offline) -- the fixtures pin the Python half of the
oracle, not Box2D.  The fixtures travel to the GPU box; /root/reference does not.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NCG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
REAL_BOX2D = "--real-box2d" in sys.argv
if REAL_BOX2D:
    OUT = os.path.join(ROOT, "tests", "golden_real")
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != HERE]      # "oracle" must resolve to the package


def _stub_path():
    """A directory holding only the stand-in modules that are needed: a module that imports for real is used for real
    (and with --real-box2d the Box2D stand-in is never offered, so a missing box2d-py is an ImportError, not a silent
    fall-back to the restatement)."""
    import importlib.util
    import tempfile
    d = tempfile.mkdtemp(prefix="ncg_refshim_")
    used = []
    for mod in ("pygame", "gymnasium", "Box2D"):
        if mod == "Box2D" and REAL_BOX2D:
            continue
        if mod != "Box2D" and importlib.util.find_spec(mod) is not None:
            continue
        os.symlink(os.path.join(HERE, "refshim", mod), os.path.join(d, mod))
        used.append(mod)
    print("stand-in modules:", used or "none", "| Box2D:", "REAL box2d-py" if REAL_BOX2D else "oracle/b2lite.h restatement")
    return d


sys.path.insert(0, _stub_path())
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def controller(obs, state):
    """A small sensor-following driver (only its recorded actions matter)."""
    s = obs[22:38]
    left = s[15] + s[14] + s[13]
    right = s[1] + s[2] + s[3]
    steer = float(np.clip(2.5 * (left - right), -1, 1))
    speed, fwd = obs[4], s[0]
    target = 0.25 + 0.55 * fwd
    tb = float(np.clip(4.0 * (target - speed), -1, 1))
    return [tb, steer]


def run(name, track, steps, policy, num_cars=1, discrete=False, reset_on_lap=False, seed=0, start_position=None, start_angle=0.0):
    from src.car_env import CarEnv
    rng = np.random.default_rng(seed)
    with quiet():
        env = CarEnv(render_mode=None, track_file=os.path.join(REF, "tracks", f"{track}.track"), num_cars=num_cars,
                     discrete_action_space=discrete, reset_on_lap=reset_on_lap, start_position=start_position, start_angle=start_angle)
        obs, info = env.reset()
    C = num_cars
    A, O, R, TE, TR, RS, LAPS, DIS, ST, REASON, ONTRACK, PERF = [], [], [], [], [], [], [], [], [], [], [], []
    obs0 = np.array(obs, dtype=np.float32).reshape(C, 38)
    cur = obs0
    for t in range(steps):
        if discrete:
            a = rng.integers(0, 5, size=C) if policy == "random" else np.array([1 if (t // 50) % 3 else 4] * C)
            act = int(a[0]) if C == 1 else a.astype(np.int64)
        else:
            if policy == "random":
                a = rng.uniform(-1, 1, size=(C, 2))
            elif policy == "drive":
                a = np.stack([rng.uniform(0.2, 1.0, size=C), rng.uniform(-0.2, 0.6, size=C)], axis=1)
            elif policy == "full":
                a = np.tile([1.0, 0.0], (C, 1))
            elif policy == "zero":
                a = np.zeros((C, 2))
            elif policy == "reverse":      # drive a little, then turn round and go backwards along the track
                a = np.tile([0.6, 0.9 if 200 < t < 420 else 0.0], (C, 1))
            else:
                a = np.array([controller(cur[c], None) for c in range(C)])
                a += rng.normal(0, 0.02, size=a.shape)
                a = np.clip(a, -1, 1)
            a = a.astype(np.float32)
            act = a[0] if C == 1 else a
        with quiet():
            o, r, te, tr, info = env.step(act)
        cur = np.array(o, dtype=np.float32).reshape(C, 38)
        A.append(np.array(a)); O.append(cur); R.append(np.array(r, dtype=np.float32).reshape(C)); TE.append(te); TR.append(tr)
        LAPS.append([info["cars"][c]["lap_timing"]["lap_count"] for c in range(C)])
        DIS.append([info["cars"][c]["disabled"] for c in range(C)])
        ONTRACK.append([info["cars"][c]["on_track"] for c in range(C)])
        PERF.append([[info["cars"][c]["performance"][k] for k in ("current_max_speed", "estimated_0_100_time", "performance_valid")]
                     for c in range(C)])
        ST.append(info["simulation_time"]); REASON.append(info["termination_reason"] or "")
        did_reset = False
        if te or tr:
            with quiet():
                o, info = env.reset()
            cur = np.array(o, dtype=np.float32).reshape(C, 38)
            did_reset = True
        RS.append(did_reset)
    with quiet():
        env.close()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"traj_{name}.npz"), track=track, num_cars=C, discrete=discrete, reset_on_lap=reset_on_lap,
                        obs0=obs0, actions=np.array(A), obs=np.array(O), reward=np.array(R), terminated=np.array(TE), truncated=np.array(TR),
                        did_reset=np.array(RS), lap_count=np.array(LAPS), disabled=np.array(DIS), on_track=np.array(ONTRACK),
                        sim_time=np.array(ST), reason=np.array(REASON), performance=np.array(PERF, dtype=np.float64),
                        start_pose=np.array([*(start_position or (0.0, 0.0)), start_angle], dtype=np.float64))
    print(f"{name}: {steps} steps, {int(np.sum(TE))} terminations, {int(np.sum(TR))} truncations, laps {np.max(LAPS)}, "
          f"disabled {int(np.sum(DIS[-1]))}, collision steps {int(np.sum(np.array(O)[:, :, 19] > 0))}, return {float(np.sum(R)):.2f}")


def dump_constants_and_tracks():
    import src.constants as C
    keep = {k: getattr(C, k) for k in dir(C) if k.isupper() and isinstance(getattr(C, k), (int, float, bool))}
    with open(os.path.join(OUT, "constants.json"), "w") as f:
        json.dump(keep, f, indent=0, sort_keys=True)
    from src.track_generator import TrackLoader
    from src.car import Car
    from src.car_physics import CarPhysics
    from src.lap_timer import LapTimer
    tracks = {}
    for fn in sorted(os.listdir(os.path.join(REF, "tracks"))):
        if not fn.endswith(".track"):
            continue
        with quiet():
            tr = TrackLoader().load_track(os.path.join(REF, "tracks", fn))
            phys = CarPhysics(Car(world=None, car_id="car_0"), tr)
            lt = LapTimer(tr)
        segs = [[s.segment_type, s.length, s.start_position[0], s.start_position[1], s.end_position[0], s.end_position[1], s.width,
                 s.curve_angle, s.curve_radius, s.curve_direction, s.start_heading, s.end_heading, s.banking_angle] for s in tr.segments]
        walls = [[b._pos0[0], b._pos0[1], b._angle0, b.fixture.shape.hx, b.fixture.shape.hy] for b in phys.world.walls]
        with open(os.path.join(REF, "tracks", fn)) as f:
            text = f.read()
        tracks[fn[:-6]] = {"width": tr.width, "total_length": tr.total_length, "segments": segs, "walls": walls,
                           "minimum_lap_distance": lt.minimum_lap_distance, "n_lines": len(text.splitlines())}
    with open(os.path.join(OUT, "tracks.json"), "w") as f:
        json.dump(tracks, f)
    print("tracks:", {k: (len(v["segments"]), len(v["walls"])) for k, v in tracks.items()})


def dump_unit_kats():
    """Pure-Python known-answer vectors straight from the reference's classes (no Box2D involved)."""
    from src.tyre_manager import TyreManager
    from src.car import Car
    rng = np.random.default_rng(7)
    kats = {"tyres": [], "rpm": [], "torque": []}
    for _ in range(64):
        tm = TyreManager()
        ff = rng.uniform(0, 3000, size=4)
        tm.set_friction_forces(dict(zip(["front_left", "front_right", "rear_left", "rear_right"], ff)))
        args = [1 / 60, float(rng.uniform(-14, 14)), float(rng.uniform(-14, 14)), float(rng.uniform(0, 110)), float(rng.uniform(0, 60))]
        seq = []
        for _ in range(5):
            tm.update(args[0], (args[1], args[2]), 0.0, args[3], args[4])
            loads, temps, wear = tm.get_observation_data()
            seq.append(loads + temps + wear + [tm.get_total_grip_coefficient()])
        kats["tyres"].append({"friction": ff.tolist(), "args": args, "out": seq})
    car = Car(world=None)
    for thr in (1.0, 0.0, 0.37, 0.9, 0.0):
        car.throttle_input = thr
        for _ in range(25):
            car._update_engine_rpm(1 / 60)
            kats["rpm"].append([thr, car.engine_rpm])
    for rpm, thr in ((1000, 1), (1800, 1), (1400, 0.5), (600, 0.2), (5600, 1.0), (9500, 0.7)):
        kats["torque"].append([rpm, thr, car._calculate_engine_torque(rpm, thr)])
    with open(os.path.join(OUT, "unit_kats.json"), "w") as f:
        json.dump(kats, f)
    print("unit kats:", {k: len(v) for k, v in kats.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if not REAL_BOX2D:                 # (pure-Python pins; the wall dump reads the stand-in's attributes)
        dump_constants_and_tracks()
        dump_unit_kats()
    run("nascar_full", "nascar", 1200, "full")
    run("nascar_zero", "nascar", 650, "zero")
    run("martinsville_drive", "martinsville", 2500, "drive", seed=1)
    run("daytona_discrete", "daytona", 1500, "random", discrete=True, seed=2)
    run("talladega_10cars", "talladega", 900, "drive", num_cars=10, seed=3)
    run("michigan_random", "michigan", 1500, "random", seed=4)
    run("martinsville_laps", "martinsville", 4200, "controller", reset_on_lap=True, seed=5)
    run("nascar2_reverse", "nascar2", 2600, "reverse", seed=6)
    run("trioval_3cars_laplimit", "trioval", 3700, "controller", num_cars=3, reset_on_lap=True, seed=7)
    run("nascar_banked_controller", "nascar_banked", 2400, "controller", seed=8)      # the one track with banking (car.py:509-566)
    # CarEnv(start_position=..., start_angle=...): a start off the racing line, pointing at the outer wall
    run("nascar_startpose", "nascar", 900, "controller", seed=9, start_position=(30.0, -4.0), start_angle=0.35)
    run("daytona_startpose_2cars", "daytona", 700, "drive", num_cars=2, seed=10, start_position=(-120.0, 6.0), start_angle=-0.2)
