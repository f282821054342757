"""ctypes binding of the CPU oracle (oracle/ncg_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(nascargymnasium_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libncg_oracle.so")
_SRC = [os.path.join(_HERE, "ncg_oracle.cpp"), os.path.join(_HERE, "b2lite.h")]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no FMA contraction, strict IEEE)."""
    os.makedirs(_BUILD, exist_ok=True)
    if not force and os.path.exists(_SO) and all(
            not os.path.exists(s) or os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRC):
        return _SO
    if not all(os.path.exists(s) for s in _SRC):
        if os.path.exists(_SO):
            return _SO
        raise FileNotFoundError("oracle sources missing")
    cmd = ["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared",
           "-o", _SO + ".tmp", _SRC[0]]
    subprocess.check_call(cmd)
    os.replace(_SO + ".tmp", _SO)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, i, d, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_float
        P = ctypes.POINTER
        L.orc_state_layout.restype = ctypes.c_char_p
        L.orc_track_parse.restype = vp
        L.orc_track_parse.argtypes = [ctypes.c_char_p]
        L.orc_track_free.argtypes = [vp]
        L.orc_track_error.restype = ctypes.c_char_p
        L.orc_track_error.argtypes = [vp]
        for n in ("orc_track_num_segs", "orc_track_num_walls"):
            getattr(L, n).restype = i
            getattr(L, n).argtypes = [vp]
        L.orc_track_total_length.restype = d
        L.orc_track_total_length.argtypes = [vp]
        L.orc_track_segs.argtypes = [vp, P(d)]
        L.orc_track_wall_lines.argtypes = [vp, P(d)]
        L.orc_env_create.restype = vp
        L.orc_env_create.argtypes = [vp, i, i]
        L.orc_env_free.argtypes = [vp]
        L.orc_env_reset.argtypes = [vp, i]
        L.orc_env_set_start.argtypes = [vp, d, d, d]
        L.orc_env_set_car_contacts.argtypes = [vp, i, d, d]
        L.orc_env_num_pairs.argtypes = [vp, i]
        L.orc_env_num_pairs.restype = i
        L.orc_env_walls.argtypes = [vp, P(f)]
        L.orc_action_continuous.argtypes = [f, f, P(f)]
        L.orc_action_discrete.argtypes = [i, P(f)]
        L.orc_action_discrete.restype = i
        L.orc_env_step.argtypes = [vp, P(f), P(f), P(f), P(i), P(i)]
        L.orc_env_observe.argtypes = [vp, P(f)]
        L.orc_env_termination_reason.argtypes = [vp]
        L.orc_env_termination_reason.restype = i
        L.orc_env_sim_time.argtypes = [vp]
        L.orc_env_sim_time.restype = d
        L.orc_env_on_track.argtypes = [vp, i]
        L.orc_env_on_track.restype = i
        L.orc_env_progress.argtypes = [vp, d, d]
        L.orc_env_progress.restype = d
        L.orc_env_impulse.argtypes = [vp, i]
        L.orc_env_impulse.restype = d
        for n in ("orc_env_num_contacts", "orc_env_num_touching"):
            getattr(L, n).argtypes = [vp, i]
            getattr(L, n).restype = i
        L.orc_env_get_state.argtypes = [vp, i, P(d)]
        L.orc_env_set_state.argtypes = [vp, i, P(d)]
        L.orc_kat_tyres.argtypes = [P(d), d, d, d, d, d, P(d)]
        L.orc_kat_rpm.argtypes = [d, d, d]
        L.orc_kat_rpm.restype = d
        L.orc_set_b2_variant.argtypes = [i]
        L.orc_get_b2_variant.restype = i
        L.orc_b2_collide.argtypes = [f] * 8 + [P(f)]
        L.orc_b2_variant_study.argtypes = [vp, ctypes.c_long, ctypes.c_ulonglong, P(ctypes.c_long)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def state_layout() -> dict:
    s = lib().orc_state_layout().decode()
    return {k: int(v) for k, v in (kv.split(":") for kv in s.strip(";").split(";"))}


TERMINATION_REASONS = (None, "all_cars_disabled", "all_active_cars_low_reward (threshold: -250.0)", "time_limit",
                       "truncated")


class OracleTrack:
    def __init__(self, text: str):
        self._h = lib().orc_track_parse(text.encode())
        err = lib().orc_track_error(self._h).decode()
        if err:
            raise ValueError(err)
        self.num_segs = lib().orc_track_num_segs(self._h)
        self.num_walls = lib().orc_track_num_walls(self._h)
        self.total_length = lib().orc_track_total_length(self._h)

    def segs(self) -> np.ndarray:
        out = np.zeros((self.num_segs, 8), dtype=np.float64)
        lib().orc_track_segs(self._h, _dp(out))
        return out

    def wall_lines(self) -> np.ndarray:
        out = np.zeros((self.num_walls, 4), dtype=np.float64)
        lib().orc_track_wall_lines(self._h, _dp(out))
        return out

    def __del__(self):
        try:
            lib().orc_track_free(self._h)
        except Exception:
            pass


class OracleEnv:
    """Sequential CPU CarEnv restatement: one env of ``num_cars`` cars."""

    def __init__(self, track_text: str, num_cars: int = 1, reset_on_lap: bool = False, discrete: bool = False,
                 start_position=(0.0, 0.0), start_angle: float = 0.0, car_contacts: bool = False, grid=(8.0, 3.0)):
        self.track = OracleTrack(track_text)
        self.num_cars = num_cars
        self.discrete = discrete
        self._h = lib().orc_env_create(self.track._h, num_cars, int(reset_on_lap))
        if tuple(start_position) != (0.0, 0.0) or start_angle != 0.0 or car_contacts:
            lib().orc_env_set_start(self._h, float(start_position[0]), float(start_position[1]), float(start_angle))
            lib().orc_env_set_car_contacts(self._h, int(car_contacts), float(grid[0]), float(grid[1]))
            lib().orc_env_reset(self._h, 1)
        self.words = state_layout()["S_WORDS"]

    def reset(self, fresh: bool = True) -> np.ndarray:
        lib().orc_env_reset(self._h, int(fresh))
        return self.observe()

    def observe(self) -> np.ndarray:
        obs = np.zeros((self.num_cars, 38), dtype=np.float32)
        lib().orc_env_observe(self._h, _fp(obs))
        return obs

    def convert_actions(self, actions) -> np.ndarray:
        """base_env.py:201-252 -> (C,3) float32 [throttle, brake, steer]."""
        out = np.zeros((self.num_cars, 3), dtype=np.float32)
        tmp = (ctypes.c_float * 3)()
        if self.discrete:
            a = np.asarray(actions).reshape(self.num_cars)
            for c in range(self.num_cars):
                if lib().orc_action_discrete(int(a[c]), tmp) != 0:
                    raise ValueError(f"Invalid discrete action: {a[c]}")
                out[c] = tmp[:]
        else:
            a = np.asarray(actions, dtype=np.float32).reshape(self.num_cars, 2)
            for c in range(self.num_cars):
                lib().orc_action_continuous(float(a[c, 0]), float(a[c, 1]), tmp)
                out[c] = tmp[:]
        return out

    def step(self, actions):
        act3 = np.ascontiguousarray(self.convert_actions(actions))
        obs = np.zeros((self.num_cars, 38), dtype=np.float32)
        rew = np.zeros(self.num_cars, dtype=np.float32)
        te, tr = ctypes.c_int(0), ctypes.c_int(0)
        lib().orc_env_step(self._h, _fp(act3), _fp(obs), _fp(rew), ctypes.byref(te), ctypes.byref(tr))
        return obs, rew, bool(te.value), bool(tr.value)

    def get_state(self, car: int = 0) -> np.ndarray:
        s = np.zeros(self.words, dtype=np.float64)
        lib().orc_env_get_state(self._h, car, _dp(s))
        return s

    def set_state(self, s: np.ndarray, car: int = 0) -> None:
        s = np.ascontiguousarray(s, dtype=np.float64)
        lib().orc_env_set_state(self._h, car, _dp(s))

    def walls(self) -> np.ndarray:
        out = np.zeros((self.track.num_walls, 10), dtype=np.float32)
        lib().orc_env_walls(self._h, _fp(out))
        return out

    @property
    def termination_reason(self):
        return TERMINATION_REASONS[lib().orc_env_termination_reason(self._h)]

    @property
    def sim_time(self) -> float:
        return lib().orc_env_sim_time(self._h)

    def on_track(self, car: int = 0) -> bool:
        return bool(lib().orc_env_on_track(self._h, car))

    def impulse(self, car: int = 0) -> float:
        return lib().orc_env_impulse(self._h, car)

    def num_pairs(self, touching_only: bool = False) -> int:
        """car-car contacts of the shared world (car_contacts=True)."""
        return lib().orc_env_num_pairs(self._h, int(touching_only))

    def num_contacts(self, car: int = 0):
        return lib().orc_env_num_contacts(self._h, car), lib().orc_env_num_touching(self._h, car)

    def __del__(self):
        try:
            lib().orc_env_free(self._h)
        except Exception:
            pass


def set_b2_variant(v: int) -> None:
    """0: b2CollidePolygons as in Box2D 2.3.1+ (default); 1: as in 2.3.0 (see oracle/b2lite.h)."""
    lib().orc_set_b2_variant(int(v))


def b2_variant_study(env: "OracleEnv", n: int, seed: int = 0) -> dict:
    c = (ctypes.c_long * 8)()
    lib().orc_b2_variant_study(env._h, int(n), int(seed), c)
    keys = ("samples", "touching", "touching_differs", "reference_face_differs", "point_count_differs", "feature_ids_differ",
            "same_features_bits_differ", "worst_point_difference_nm")
    return dict(zip(keys, [int(x) for x in c]))


def b2_collide(car_pose, wall_pose, wall_half) -> np.ndarray:
    out = np.zeros(12, dtype=np.float32)
    lib().orc_b2_collide(*[float(x) for x in car_pose], *[float(x) for x in wall_pose], float(wall_half[0]), float(wall_half[1]), _fp(out))
    return out


def kat_tyres(friction4, dt, along, alat, speed, slip):
    f = np.asarray(friction4, dtype=np.float64)
    out = np.zeros(13, dtype=np.float64)
    lib().orc_kat_tyres(_dp(f), dt, along, alat, speed, slip, _dp(out))
    return out


def kat_rpm(rpm, throttle, dt=1.0 / 60.0):
    return lib().orc_kat_rpm(rpm, throttle, dt)
