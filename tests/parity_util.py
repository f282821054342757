"""Test helpers: oracle state <-> engine record conversion, tolerances, the hostcheck build.

Everything here is test infrastructure (it imports oracle/)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from nascargymnasium_b200 import layout as L
from nascargymnasium_b200 import track as T
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, F = L.R, L.F
S = O.state_layout()
W = L.RECORD_WORDS

# t(n): the reference's simulation_time after n additions of 1/60 in float64 (car_env.py:573)
TIMES = np.concatenate([[0.0], np.add.accumulate(np.full(30000, 1.0 / 60.0, dtype=np.float64))])


def steps_of_time(t: float) -> int:
    n = int(np.searchsorted(TIMES, t - 1e-9))
    assert abs(TIMES[n] - t) < 1e-9, (t, n, TIMES[n])
    return n


def _u(rec, idx):
    return int(rec.view(np.uint32)[idx])


def _setu(rec, idx, val):
    rec.view(np.uint32)[idx] = np.uint32(val)


def oracle_to_record(s: np.ndarray, track_id: int = 0) -> np.ndarray:
    """Oracle flat state (float64) -> engine record (float32 words).  Raises if a cap is exceeded."""
    rec = np.zeros(W, dtype=np.float32)
    for a, b in (("S_X", "NCG_R_X"), ("S_Y", "NCG_R_Y"), ("S_A", "NCG_R_ANGLE"), ("S_VX", "NCG_R_VX"), ("S_VY", "NCG_R_VY"),
                 ("S_W", "NCG_R_OMEGA"), ("S_SLEEP", "NCG_R_SLEEP"), ("S_FLX", "NCG_R_FAT_LX"), ("S_FLY", "NCG_R_FAT_LY"),
                 ("S_FUX", "NCG_R_FAT_UX"), ("S_FUY", "NCG_R_FAT_UY"), ("S_INVDT0", "NCG_R_INV_DT0"), ("S_IMPULSE", "NCG_R_IMPULSE"),
                 ("S_RPM", "NCG_R_RPM"), ("S_PVX", "NCG_R_PREV_VX"), ("S_PVY", "NCG_R_PREV_VY"), ("S_SLIP", "NCG_R_SLIP"),
                 ("S_FLAT", "NCG_R_FLAT"), ("S_BANK", "NCG_R_BANK"), ("S_CUMIMP", "NCG_R_CUM_IMPACT"), ("S_STUCKX", "NCG_R_STUCK_X"),
                 ("S_STUCKY", "NCG_R_STUCK_Y"), ("S_BACK", "NCG_R_BACK"), ("S_BACKPREV", "NCG_R_BACK_PREV"),
                 ("S_PPREV", "NCG_R_PROGRESS_PREV"), ("S_PREVX", "NCG_R_PREV_X"), ("S_PREVY", "NCG_R_PREV_Y"),
                 ("S_CUMREW", "NCG_R_CUM_REWARD"), ("S_LAST", "NCG_R_LAST_LAP"), ("S_BEST", "NCG_R_BEST_LAP"), ("S_ODO", "NCG_R_ODO"),
                 ("S_LX", "NCG_R_LAP_X"), ("S_LY", "NCG_R_LAP_Y")):
        rec[R[b]] = np.float32(s[S[a]])
    for k in range(4):
        rec[R["NCG_R_TYRE_TEMP"] + k] = s[S["S_TT"] + k]
        rec[R["NCG_R_TYRE_WEAR"] + k] = s[S["S_TW"] + k]
        rec[R["NCG_R_TYRE_LOAD"] + k] = s[S["S_TL"] + k]
    n = int(s[S["S_ACCN"]])
    _setu(rec, R["NCG_R_ACC_N"], n)
    rec[R["NCG_R_ACC"]:R["NCG_R_ACC"] + 2 * n] = s[S["S_ACC"]:S["S_ACC"] + 2 * n]
    fl = 0
    pm = int(s[S["S_PROXYMOVED"]])
    fl |= F["NCG_F_AWAKE"] if s[S["S_AWAKE"]] else 0
    fl |= F["NCG_F_PROXY_MOVED"] if pm & 1 else 0
    fl |= F["NCG_F_NEW_FIXTURE"] if pm & 2 else 0
    fl |= F["NCG_F_HAS_KEY"] if s[S["S_HASKEY"]] else 0
    fl |= F["NCG_F_DISABLED"] if s[S["S_DISABLED"]] else 0
    fl |= F["NCG_F_FIRST_STEP"] if s[S["S_FIRST"]] else 0
    fl |= F["NCG_F_STUCK_POS"] if s[S["S_STUCKVALID"]] else 0
    fl |= F["NCG_F_BACK_ACTIVE"] if s[S["S_BACKACTIVE"]] else 0
    fl |= F["NCG_F_CROSSED"] if s[S["S_CROSSED"]] else 0
    fl |= F["NCG_F_HAS_LAST"] if s[S["S_HASLAST"]] else 0
    fl |= F["NCG_F_HAS_BEST"] if s[S["S_HASBEST"]] else 0
    fl |= F["NCG_F_HAS_POS"] if s[S["S_HASPOS"]] else 0
    _setu(rec, R["NCG_R_FLAGS"], fl)
    # contacts
    nc = int(s[S["S_NCONTACT"]])
    na = int(s[S["S_NACTIVE"]])
    if s[S["S_OVERFLOW"]] or nc > L.MAX_CONTACTS or na > L.MAX_ACTIVE:
        raise OverflowError("oracle state exceeds the engine's per-car caps")
    tmask, pcw, k = 0, 0, 0
    walls = [0] * 6
    for i in range(nc):
        o = s[S["S_CONTACTS"] + 9 * i: S["S_CONTACTS"] + 9 * i + 9]
        walls[i >> 1] |= (int(o[0]) & 0xFFFF) << ((i & 1) * 16)
        if o[1]:
            if k >= L.MAX_TOUCHING:
                raise OverflowError("too many touching contacts")
            pc = int(o[2])
            tmask |= 1 << i
            pcw |= pc << (2 * k)
            M = R["NCG_R_MANIFOLD"] + 6 * k
            _setu(rec, M, int(o[3]))
            _setu(rec, M + 1, int(o[4]) if pc > 1 else 0)
            rec[M + 2], rec[M + 3] = o[5], o[6]
            if pc > 1:
                rec[M + 4], rec[M + 5] = o[7], o[8]
            k += 1
    for i in range(6):
        _setu(rec, R["NCG_R_CONTACT_WALL"] + i, walls[i])
    _setu(rec, R["NCG_R_NCONTACT"], nc | (na << 8) | (tmask << 16))
    _setu(rec, R["NCG_R_MANIFOLD_PC"], pcw)
    for i in range(na):
        A = R["NCG_R_ACTIVE"] + 3 * i
        _setu(rec, A, int(s[S["S_ACTIVE"] + 3 * i]))
        rec[A + 1], rec[A + 2] = s[S["S_ACTIVE"] + 3 * i + 1], s[S["S_ACTIVE"] + 3 * i + 2]
    _setu(rec, R["NCG_R_STUCK_STEPS"], steps_of_time(s[S["S_STUCKT"]]))
    _setu(rec, R["NCG_R_LAP_START"], steps_of_time(s[S["S_LAPSTART"]]) if s[S["S_TIMING"]] else 0)
    _setu(rec, R["NCG_R_LAP_COUNT"], int(s[S["S_LAPCOUNT"]]))
    _setu(rec, R["NCG_R_STEP"], steps_of_time(s[S["S_SIMTIME"]]))
    _setu(rec, R["NCG_R_TRACK"], track_id)
    return rec


def record_to_oracle(rec: np.ndarray) -> np.ndarray:
    """Engine record -> oracle flat state (every value float32-representable)."""
    rec = np.ascontiguousarray(rec, dtype=np.float32)
    s = np.zeros(S["S_WORDS"], dtype=np.float64)
    for a, b in (("S_X", "NCG_R_X"), ("S_Y", "NCG_R_Y"), ("S_A", "NCG_R_ANGLE"), ("S_VX", "NCG_R_VX"), ("S_VY", "NCG_R_VY"),
                 ("S_W", "NCG_R_OMEGA"), ("S_SLEEP", "NCG_R_SLEEP"), ("S_FLX", "NCG_R_FAT_LX"), ("S_FLY", "NCG_R_FAT_LY"),
                 ("S_FUX", "NCG_R_FAT_UX"), ("S_FUY", "NCG_R_FAT_UY"), ("S_INVDT0", "NCG_R_INV_DT0"), ("S_IMPULSE", "NCG_R_IMPULSE"),
                 ("S_RPM", "NCG_R_RPM"), ("S_PVX", "NCG_R_PREV_VX"), ("S_PVY", "NCG_R_PREV_VY"), ("S_SLIP", "NCG_R_SLIP"),
                 ("S_FLAT", "NCG_R_FLAT"), ("S_BANK", "NCG_R_BANK"), ("S_CUMIMP", "NCG_R_CUM_IMPACT"), ("S_STUCKX", "NCG_R_STUCK_X"),
                 ("S_STUCKY", "NCG_R_STUCK_Y"), ("S_BACK", "NCG_R_BACK"), ("S_BACKPREV", "NCG_R_BACK_PREV"),
                 ("S_PPREV", "NCG_R_PROGRESS_PREV"), ("S_PREVX", "NCG_R_PREV_X"), ("S_PREVY", "NCG_R_PREV_Y"),
                 ("S_CUMREW", "NCG_R_CUM_REWARD"), ("S_LAST", "NCG_R_LAST_LAP"), ("S_BEST", "NCG_R_BEST_LAP"), ("S_ODO", "NCG_R_ODO"),
                 ("S_LX", "NCG_R_LAP_X"), ("S_LY", "NCG_R_LAP_Y")):
        s[S[a]] = float(rec[R[b]])
    for k in range(4):
        s[S["S_TT"] + k] = rec[R["NCG_R_TYRE_TEMP"] + k]
        s[S["S_TW"] + k] = rec[R["NCG_R_TYRE_WEAR"] + k]
        s[S["S_TL"] + k] = rec[R["NCG_R_TYRE_LOAD"] + k]
    n = _u(rec, R["NCG_R_ACC_N"])
    s[S["S_ACCN"]] = n
    s[S["S_ACC"]:S["S_ACC"] + 2 * n] = rec[R["NCG_R_ACC"]:R["NCG_R_ACC"] + 2 * n]
    fl = _u(rec, R["NCG_R_FLAGS"])
    s[S["S_AWAKE"]] = bool(fl & F["NCG_F_AWAKE"])
    s[S["S_PROXYMOVED"]] = (1 if fl & F["NCG_F_PROXY_MOVED"] else 0) + (2 if fl & F["NCG_F_NEW_FIXTURE"] else 0)
    s[S["S_HASKEY"]] = bool(fl & F["NCG_F_HAS_KEY"])
    s[S["S_DISABLED"]] = bool(fl & F["NCG_F_DISABLED"])
    s[S["S_FIRST"]] = bool(fl & F["NCG_F_FIRST_STEP"])
    s[S["S_STUCKVALID"]] = bool(fl & F["NCG_F_STUCK_POS"])
    s[S["S_BACKACTIVE"]] = bool(fl & F["NCG_F_BACK_ACTIVE"])
    s[S["S_CROSSED"]] = s[S["S_TIMING"]] = bool(fl & F["NCG_F_CROSSED"])
    s[S["S_HASLAST"]] = bool(fl & F["NCG_F_HAS_LAST"])
    s[S["S_HASBEST"]] = bool(fl & F["NCG_F_HAS_BEST"])
    s[S["S_HASPOS"]] = bool(fl & F["NCG_F_HAS_POS"])
    ncw = _u(rec, R["NCG_R_NCONTACT"])
    nc, na, tmask = ncw & 255, (ncw >> 8) & 255, (ncw >> 16) & 0xFFF
    pcw = _u(rec, R["NCG_R_MANIFOLD_PC"])
    s[S["S_NCONTACT"]], s[S["S_NACTIVE"]] = nc, na
    k = 0
    for i in range(nc):
        ww = _u(rec, R["NCG_R_CONTACT_WALL"] + (i >> 1))
        o = S["S_CONTACTS"] + 9 * i
        s[o] = (ww >> 16) if (i & 1) else (ww & 0xFFFF)
        if (tmask >> i) & 1:
            M = R["NCG_R_MANIFOLD"] + 6 * k
            s[o + 1] = 1
            s[o + 2] = (pcw >> (2 * k)) & 3
            s[o + 3], s[o + 4] = _u(rec, M), _u(rec, M + 1)
            s[o + 5], s[o + 6], s[o + 7], s[o + 8] = rec[M + 2], rec[M + 3], rec[M + 4], rec[M + 5]
            k += 1
    for i in range(na):
        A = R["NCG_R_ACTIVE"] + 3 * i
        s[S["S_ACTIVE"] + 3 * i] = _u(rec, A)
        s[S["S_ACTIVE"] + 3 * i + 1], s[S["S_ACTIVE"] + 3 * i + 2] = rec[A + 1], rec[A + 2]
    s[S["S_STUCKT"]] = TIMES[_u(rec, R["NCG_R_STUCK_STEPS"])]
    s[S["S_LAPSTART"]] = TIMES[_u(rec, R["NCG_R_LAP_START"])]
    s[S["S_LAPCOUNT"]] = s[S["S_PREVLAPS"]] = _u(rec, R["NCG_R_LAP_COUNT"])
    s[S["S_SIMTIME"]] = TIMES[_u(rec, R["NCG_R_STEP"])]
    return s


# float32 words compared with a relative/absolute tolerance; u32 words compared exactly
FLOAT_FIELDS = {
    # name: (count, rtol, atol)   north_star: position, velocity and tyre state within 1e-4 relative
    "NCG_R_X": (2, 1e-4, 1e-4), "NCG_R_ANGLE": (1, 1e-4, 1e-5), "NCG_R_VX": (2, 1e-4, 1e-4), "NCG_R_OMEGA": (1, 1e-4, 1e-4),
    "NCG_R_SLEEP": (1, 1e-5, 1e-6), "NCG_R_FAT_LX": (4, 1e-4, 1e-3), "NCG_R_INV_DT0": (1, 1e-6, 0), "NCG_R_IMPULSE": (1, 1e-4, 1e-2),
    "NCG_R_RPM": (1, 1e-5, 1e-3), "NCG_R_PREV_VX": (2, 1e-4, 1e-4), "NCG_R_SLIP": (1, 1e-3, 1e-2), "NCG_R_FLAT": (1, 1e-4, 1e-1),
    "NCG_R_BANK": (1, 0, 0), "NCG_R_TYRE_TEMP": (4, 1e-4, 1e-4), "NCG_R_TYRE_WEAR": (4, 1e-4, 1e-6), "NCG_R_TYRE_LOAD": (4, 1e-4, 1e-2),
    "NCG_R_CUM_IMPACT": (1, 1e-4, 1e-1), "NCG_R_STUCK_X": (2, 1e-4, 1e-4), "NCG_R_BACK": (2, 1e-3, 1e-3),
    "NCG_R_PROGRESS_PREV": (1, 1e-4, 1e-3), "NCG_R_PREV_X": (2, 1e-4, 1e-4), "NCG_R_CUM_REWARD": (1, 1e-4, 1e-4),
    "NCG_R_LAST_LAP": (2, 1e-5, 1e-5), "NCG_R_ODO": (1, 1e-4, 1e-3), "NCG_R_LAP_X": (2, 1e-4, 1e-4),
}
UINT_FIELDS = ("NCG_R_ACC_N", "NCG_R_STUCK_STEPS", "NCG_R_LAP_START", "NCG_R_LAP_COUNT", "NCG_R_STEP")
FLAG_MASK = sum(F[k] for k in ("NCG_F_AWAKE", "NCG_F_HAS_KEY", "NCG_F_DISABLED", "NCG_F_FIRST_STEP", "NCG_F_STUCK_POS",
                               "NCG_F_BACK_ACTIVE", "NCG_F_CROSSED", "NCG_F_HAS_LAST", "NCG_F_HAS_BEST", "NCG_F_HAS_POS"))


def compare_records(got: np.ndarray, want: np.ndarray, contact_fields: bool = True, touching: bool = False):
    """Returns a list of (field, got, want) mismatches between an engine record and the oracle's.

    `touching`: the step solved contact constraints.  Box2D's 2-point block solver accepts effective-mass matrices with
    condition numbers up to 1000 (b2ContactSolver: k_maxConditionNumber), which amplifies float32 rounding of its inputs
    by that factor into the angular impulse, so on such steps the angular velocity is compared at 3e-4 rad/s absolute
    and the stored manifold impulses at 0.1 N s + 4e-4 relative.  Measured worst cases over 4332 contact-solving steps on five
    tracks (tools/contact_worst_case.py -> profiles/r02_contact_worst_case.json): omega 1.7e-4 rad/s, angle 1.6e-4 rad,
    velocity 1.1e-5 relative, position 7.6e-7 relative, manifold impulses 0.065 N s / 1.9e-4 relative, cumulative impact
    2.3e-5 relative, listener impulse identical."""
    bad = []
    for name, (cnt, rtol, atol) in FLOAT_FIELDS.items():
        if touching and name == "NCG_R_OMEGA":
            atol = 3e-4
        if touching and name == "NCG_R_ANGLE":
            atol = 3e-4
        a, b = got[R[name]:R[name] + cnt].astype(np.float64), want[R[name]:R[name] + cnt].astype(np.float64)
        if not np.all(np.abs(a - b) <= atol + rtol * np.abs(b)):
            bad.append((name, a.copy(), b.copy()))
    n = _u(want, R["NCG_R_ACC_N"])
    a, b = got[R["NCG_R_ACC"]:R["NCG_R_ACC"] + 2 * n].astype(np.float64), want[R["NCG_R_ACC"]:R["NCG_R_ACC"] + 2 * n].astype(np.float64)
    if not np.all(np.abs(a - b) <= 2e-3 + 1e-4 * np.abs(b)):
        bad.append(("NCG_R_ACC", a, b))
    for name in UINT_FIELDS:
        if _u(got, R[name]) != _u(want, R[name]):
            bad.append((name, _u(got, R[name]), _u(want, R[name])))
    if (_u(got, R["NCG_R_FLAGS"]) ^ _u(want, R["NCG_R_FLAGS"])) & FLAG_MASK:
        bad.append(("NCG_R_FLAGS", hex(_u(got, R["NCG_R_FLAGS"])), hex(_u(want, R["NCG_R_FLAGS"]))))
    if contact_fields:
        if _u(got, R["NCG_R_NCONTACT"]) != _u(want, R["NCG_R_NCONTACT"]):
            bad.append(("NCG_R_NCONTACT", hex(_u(got, R["NCG_R_NCONTACT"])), hex(_u(want, R["NCG_R_NCONTACT"]))))
        else:
            for i in range(6):
                if _u(got, R["NCG_R_CONTACT_WALL"] + i) != _u(want, R["NCG_R_CONTACT_WALL"] + i):
                    bad.append(("NCG_R_CONTACT_WALL", i, hex(_u(got, R["NCG_R_CONTACT_WALL"] + i)), hex(_u(want, R["NCG_R_CONTACT_WALL"] + i))))
            if _u(got, R["NCG_R_MANIFOLD_PC"]) != _u(want, R["NCG_R_MANIFOLD_PC"]):
                bad.append(("NCG_R_MANIFOLD_PC", _u(got, R["NCG_R_MANIFOLD_PC"]), _u(want, R["NCG_R_MANIFOLD_PC"])))
            ncw = _u(want, R["NCG_R_NCONTACT"])
            nt = bin((ncw >> 16) & 0xFFF).count("1")
            for k in range(nt):
                M = R["NCG_R_MANIFOLD"] + 6 * k
                if _u(got, M) != _u(want, M) or _u(got, M + 1) != _u(want, M + 1):
                    bad.append(("manifold keys", k))
                a, b = got[M + 2:M + 6].astype(np.float64), want[M + 2:M + 6].astype(np.float64)
                if not np.all(np.abs(a - b) <= 0.1 + 4e-4 * np.abs(b)):
                    bad.append(("manifold impulses", a, b))
            na = (ncw >> 8) & 255
            for i in range(na):
                A = R["NCG_R_ACTIVE"] + 3 * i
                if _u(got, A) != _u(want, A) or not np.allclose(got[A + 1:A + 3], want[A + 1:A + 3], atol=1e-4):
                    bad.append(("active collision", i))
    return bad


# ----------------------------------------------------------------------------- hostcheck (CPU compile of the device code)
_HC = {}


def hostcheck(no_toi_shortcut: bool = False):
    """The host compile of the device code (tests/hostcheck).  no_toi_shortcut: the same source with the TOI early-out of
    w_solve_toi compiled out (-DNCG_NO_TOI_SHORTCUT), to prove the early-out changes nothing."""
    key = bool(no_toi_shortcut)
    if key not in _HC:
        src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
        out = os.path.join(ROOT, "tests", "hostcheck", "_build", "libncg_hostcheck_notoi.so" if key else "libncg_hostcheck.so")
        deps = [src] + [os.path.join(ROOT, "nascargymnasium_b200", "csrc", f) for f in ("ncg_car.cuh", "ncg_b2.cuh", "ncg_defs.cuh")] + \
               [os.path.join(ROOT, "include", "ncg_b200.h")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                                   "-Wno-unknown-pragmas"] + (["-DNCG_NO_TOI_SHORTCUT"] if key else []) + ["-o", out, src])
        lib = ctypes.CDLL(out)
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        lib.hc_env_step.argtypes = [fp, fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, fp, fp, ip, ip, ip,
                                    ctypes.POINTER(ctypes.c_ulonglong)]
        lib.hc_env_reset.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        lib.hc_sensors_brute.argtypes = [fp, ctypes.c_float, ctypes.c_float, ctypes.c_float, fp]
        lib.hc_sensors_grid.argtypes = [fp, ctypes.c_float, ctypes.c_float, ctypes.c_float, fp, ctypes.POINTER(ctypes.c_uint)]
        lib.hc_sensors_multi_mismatches.argtypes = [fp, ctypes.c_float, ctypes.c_float, ctypes.c_float]
        lib.hc_sensors_multi_mismatches.restype = ctypes.c_int
        lib.hc_nearest_segment.argtypes = [fp, ctypes.c_float, ctypes.c_float, ctypes.c_int, fp]
        lib.hc_on_track.argtypes = [fp, ctypes.c_float, ctypes.c_float]
        lib.hc_synthetic_action.argtypes = [ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, ctypes.c_int, ctypes.c_int, fp]
        lib.hc_sweep_face_bound.argtypes = [fp, fp]
        lib.hc_sweep_face_bound.restype = ctypes.c_float
        lib.hc_env_step_cc.argtypes = [fp, fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, fp, fp, fp, ip, ip, ip,
                                       ctypes.POINTER(ctypes.c_ulonglong)]
        lib.hc_env_reset_cc.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, fp, fp]
        _HC[key] = lib
    return _HC[key]


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class HostCheckEnv:
    """One env of C cars stepped by the host compile of the device code."""

    def __init__(self, track_name: str, num_cars: int = 1, reset_on_lap: bool = False, contacts: bool = True,
                 no_toi_shortcut: bool = False, car_contacts: bool = False, grid=(8.0, 3.0)):
        self.lib = hostcheck(no_toi_shortcut)
        self.car_contacts, self.grid = car_contacts, grid
        self.pairs = np.zeros(self.lib.hc_cc_stride(), dtype=np.float32)
        self.table = T.get_track_table(track_name)
        self.blob = np.ascontiguousarray(self.table.blob)
        self.C = num_cars
        self.reset_on_lap = reset_on_lap
        self.contacts = contacts
        self.records = np.zeros((num_cars, W), dtype=np.float32)
        self.counters = (ctypes.c_ulonglong * 5)()

    def reset(self, fresh=True):
        obs = np.zeros((self.C, 38), dtype=np.float32)
        if self.car_contacts:
            self.lib.hc_env_reset_cc(_fp(self.blob), _fp(self.records), self.C, int(fresh), 0, self.grid[0], self.grid[1], _fp(self.pairs), _fp(obs))
        else:
            self.lib.hc_env_reset(_fp(self.blob), _fp(self.records), self.C, int(fresh), 0, _fp(obs))
        return obs

    def step(self, act3):
        act3 = np.ascontiguousarray(act3, dtype=np.float32).reshape(self.C, 3)
        obs = np.zeros((self.C, 38), dtype=np.float32)
        rew = np.zeros(self.C, dtype=np.float32)
        te, tr, why = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if self.car_contacts:
            self.lib.hc_env_step_cc(_fp(self.blob), _fp(self.records), self.C, _fp(act3), int(self.contacts), int(self.reset_on_lap),
                                    _fp(self.pairs), _fp(obs), _fp(rew), ctypes.byref(te), ctypes.byref(tr), ctypes.byref(why), self.counters)
        else:
            self.lib.hc_env_step(_fp(self.blob), _fp(self.records), self.C, _fp(act3), int(self.contacts), int(self.reset_on_lap),
                                 _fp(obs), _fp(rew), ctypes.byref(te), ctypes.byref(tr), ctypes.byref(why), self.counters)
        return obs, rew, bool(te.value), bool(tr.value), why.value


def philox_actions_np(seed: int, car: int, step: int, mode: int = 0, discrete: bool = False) -> np.ndarray:
    out = np.zeros(3, dtype=np.float32)
    hostcheck().hc_synthetic_action(seed, car, step, mode, int(discrete), _fp(out))
    return out


# ----------------------------------------------------------------------------- teacher-forced cases
def policy_actions(kind: str, rng, discrete: bool):
    if discrete:
        return int(rng.integers(0, 5))
    if kind == "drive":
        return [float(rng.uniform(0.2, 1.0)), float(rng.uniform(-0.2, 0.6))]
    if kind == "full":
        return [1.0, 0.0]
    return [float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1))]


def collect_cases(track: str, n_cases: int, kind: str = "drive", seed: int = 0, discrete: bool = False, every: int = 7,
                  max_steps: int = 200000, b2_variant: int = 0):
    """Run the oracle (one car) with a scripted policy and sample teacher-forcing cases along the way.

    Returns (records[n,128] float32, act3[n,3] float32, raw actions list, expected dict) where expected holds
    the oracle's results of stepping each sampled state once: records, obs, reward, terminated, truncated."""
    rng = np.random.default_rng(seed)
    O.set_b2_variant(b2_variant)
    try:
        return _collect_cases(track, n_cases, kind, rng, discrete, every, max_steps)
    finally:
        O.set_b2_variant(0)


def _collect_cases(track, n_cases, kind, rng, discrete, every, max_steps):
    env = O.OracleEnv(T.builtin_track_text(track), discrete=discrete)
    probe = O.OracleEnv(T.builtin_track_text(track), discrete=discrete)
    env.reset()
    recs, acts3, raws, exp_rec, exp_obs, exp_rew, exp_te, exp_tr, touching = [], [], [], [], [], [], [], [], []
    step = 0
    while len(recs) < n_cases and step < max_steps:
        a = policy_actions(kind, rng, discrete)
        if step % every == 0:
            try:
                rec = oracle_to_record(env.get_state())
            except OverflowError:
                rec = None
            if rec is not None:
                probe.set_state(record_to_oracle(rec))
                ob, rw, te, tr = probe.step([a])
                try:
                    want = oracle_to_record(probe.get_state())
                except OverflowError:
                    want = None
                if want is not None:
                    recs.append(rec); acts3.append(probe.convert_actions([a])[0]); raws.append(a)
                    exp_rec.append(want); exp_obs.append(ob[0]); exp_rew.append(rw[0]); exp_te.append(te); exp_tr.append(tr)
                    touching.append(probe.num_contacts()[1])
        _, _, te, tr = env.step([a])
        if te or tr:
            env.reset(fresh=False)
        step += 1
    exp = dict(records=np.array(exp_rec), obs=np.array(exp_obs), reward=np.array(exp_rew), terminated=np.array(exp_te),
               truncated=np.array(exp_tr), touching=np.array(touching))
    return np.array(recs, dtype=np.float32), np.array(acts3, dtype=np.float32), raws, exp


def check_cases(got_records, got_obs, got_reward, got_te, got_tr, exp, label=""):
    """Compares engine results with the oracle's.  Returns (n_bad, report)."""
    n = len(exp["reward"])
    bad, lines = 0, []
    for i in range(n):
        touching = bool(exp["touching"][i] > 0)
        b = compare_records(got_records[i], exp["records"][i], touching=touching)
        d22 = np.abs(got_obs[i, :22] - exp["obs"][i, :22])
        if touching:
            d22[6] = max(0.0, d22[6] - 1e-4)          # obs[6] = omega / 10: see compare_records
        dobs = float(d22.max())
        dsens = float(np.abs(got_obs[i, 22:] - exp["obs"][i, 22:]).max())
        drew = abs(float(got_reward[i]) - float(exp["reward"][i]))
        flags_ok = bool(got_te[i]) == bool(exp["terminated"][i]) and bool(got_tr[i]) == bool(exp["truncated"][i])
        if b or dobs > 1e-4 or dsens > 1e-3 or drew > 1e-4 + 1e-4 * abs(float(exp["reward"][i])) or not flags_ok:
            bad += 1
            if len(lines) < 10:
                lines.append(f"{label} case {i}: touching={exp['touching'][i]} dobs={dobs:.2e} dsens={dsens:.2e} drew={drew:.2e} "
                             f"flags_ok={flags_ok} fields={[x[0] for x in b]}")
    return bad, "\n".join(lines)
