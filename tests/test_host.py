"""Host-side logic: track ingestion vs the reference's own loader output (tests/golden/tracks.json), constants shared
with the device, step-count thresholds, spaces, and that the C-ABI library loads and exports every declared symbol."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from nascargymnasium_b200 import constants as K
from nascargymnasium_b200 import layout as L
from nascargymnasium_b200 import spaces as S
from nascargymnasium_b200 import track as T
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "tracks.json")) as f:
    REF_TRACKS = json.load(f)


def test_builtin_tracks_are_the_reference_track_files():
    assert sorted(REF_TRACKS) == sorted(T.BUILTIN_TRACK_NAMES)
    if os.path.isdir("/root/reference/tracks"):        # build container only
        for name in T.BUILTIN_TRACK_NAMES:
            with open(f"/root/reference/tracks/{name}.track") as f:
                ref = T.parse_track_text(f.read(), name)
            mine = T.load_track(name)
            assert [(s.kind, s.length, s.start, s.end, s.width, s.banking) for s in ref.segments] == \
                   [(s.kind, s.length, s.start, s.end, s.width, s.banking) for s in mine.segments]


@pytest.mark.parametrize("name", sorted(REF_TRACKS))
def test_segments_and_walls_match_reference_loader(name):
    ref = REF_TRACKS[name]
    tr = T.load_track(name)
    assert tr.total_length == ref["total_length"] and tr.width == ref["width"]
    assert len(tr.segments) == len(ref["segments"])
    for s, r in zip(tr.segments, ref["segments"]):
        assert s.kind == r[0]
        got = [s.length, s.start[0], s.start[1], s.end[0], s.end[1], s.width, s.curve_angle, s.curve_radius, s.start_heading,
               s.end_heading, s.banking]
        want = [r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[10], r[11], r[12]]
        assert got == want                     # bit-for-bit: same float64 operations in the same order
    # walls as the reference hands them to Box2D: centre, angle, half length, half thickness
    lines = T.wall_lines(tr)
    assert len(lines) == len(ref["walls"])
    boxes = T.wall_boxes(tr)
    for (x1, y1, x2, y2), w, b in zip(lines, ref["walls"], boxes):
        assert (x1 + x2) / 2 == w[0] and (y1 + y2) / 2 == w[1]
        assert np.float32(np.arctan2(y2 - y1, x2 - x1)) == np.float32(w[2])
        assert b[4] == np.float32(w[3]) and b[5] == np.float32(w[4]) and b[6] == np.float32(w[2])
    # the oracle parses the same text independently
    ot = O.OracleTrack(T.builtin_track_text(name))
    assert ot.total_length == ref["total_length"]
    assert np.array_equal(ot.wall_lines(), lines)
    assert ref["minimum_lap_distance"] == pytest.approx(tr.total_length * K.MIN_LAP_DISTANCE_FRACTION, rel=1e-15)


@pytest.mark.parametrize("name", ["nascar", "talladega"])
def test_wall_table_equals_what_box2d_stores(name):
    env = O.OracleEnv(T.builtin_track_text(name))
    w = env.walls()                                       # px,py,c,s,hx,hy, fat lx,ly,ux,uy from oracle/b2lite.h
    tab = T.get_track_table(name)
    assert np.array_equal(tab.boxes[:, :6], w[:, :6])
    assert np.array_equal(tab.fat_aabb, w[:, 6:10])


def test_track_parser_errors():
    with pytest.raises(ValueError):
        T.parse_track_text("WIDTH 10\nGRID\nLOOP 3\n")
    with pytest.raises(ValueError):
        T.parse_track_text("LEFT 400 10\n")
    with pytest.raises(ValueError):
        T.parse_track_text("STRAIGHT 10 60\n")
    with pytest.raises(FileNotFoundError):
        T.load_track("tracks/does_not_exist.track")
    t = T.parse_track_text("# c\nwidth 12 # inline\ngrid\nstartline\nleft 90 50 10\n")
    assert t.width == 12 and [s.kind for s in t.segments] == ["GRID", "STARTLINE", "CURVE"]


def test_grid_lists_every_wall_in_every_cell_it_touches():
    for name in ("martinsville", "daytona"):
        tab = T.get_track_table(name)
        cell = float(tab.blob[T.H_CELL])
        nx, ny = tab.grid_dims
        x0, y0 = tab.grid_origin
        cs = T._wall_corners(tab.boxes)
        hdr = tab.blob.view(np.uint32)
        cells = hdr[hdr[T.H_OFF_CELLS]: hdr[T.H_OFF_CELLS] + nx * ny]
        items = tab.blob.view(np.uint16)[2 * hdr[T.H_OFF_ITEMS]:]
        for w in range(0, tab.n_walls, 37):
            for cx, cy in cs[w]:
                ix, iy = int((cx - x0) // cell), int((cy - y0) // cell)
                c = iy * nx + ix
                assert w in tab.cell_items[tab.cell_start[c]:tab.cell_start[c + 1]]
                # device layout: u32 per cell = first block | n_blocks << 16, blocks of four u16 wall indices
                first, nblk = int(cells[c] & 0xFFFF), int(cells[c] >> 16)
                dev = items[4 * first: 4 * (first + nblk)]
                assert w in dev and set(dev.tolist()) == set(tab.cell_items[tab.cell_start[c]:tab.cell_start[c + 1]].tolist())
        assert len(tab.blob) % 4 == 0 and tab.blob.view(np.int32)[T.H_STAGE_WORDS] % 4 == 0
        assert tab.blob.nbytes <= T.MAX_TABLE_BYTES


def test_cell_size_is_chosen_per_track():
    """Half-mile tracks (1 m chords) get a finer grid than the superspeedways; the choice is deterministic."""
    assert float(T.get_track_table("martinsville").blob[T.H_CELL]) < float(T.get_track_table("daytona").blob[T.H_CELL])
    a = T.build_best_track_table(T.load_track("nascar2"))
    assert np.array_equal(a.blob.view(np.uint32), T.get_track_table("nascar2").blob.view(np.uint32))


def test_step_count_thresholds_are_the_float64_clock():
    """The reference adds 1/60 to a float64 clock; the engine counts steps."""
    t = np.concatenate([[0.0], np.add.accumulate(np.full(11000, 1.0 / 60.0))])
    assert int(np.argmax(t > K.STUCK_TIME)) == K.STUCK_STEPS == 600
    assert int(np.argmax(t > K.STUCK_EXTENDED_TIME)) == K.STUCK_EXTENDED_STEPS == 900
    assert int(np.argmax(t > K.TERMINATION_MAX_TIME)) == K.TERMINATION_STEPS == 3601
    assert int(np.argmax(t > K.TRUNCATION_MAX_TIME)) == K.TRUNCATION_STEPS == 10800


def test_device_constants_match_host_constants():
    with open(os.path.join(ROOT, "nascargymnasium_b200", "csrc", "ncg_defs.cuh")) as f:
        txt = f.read()
    defs = dict(re.findall(r"#define\s+(NCG_[A-Z0-9_]+)\s+\(?(-?[0-9.]+)f?\)?\s", txt))
    want = {"NCG_CAR_MASS": K.CAR_MASS, "NCG_CAR_HALF_LENGTH": K.CAR_LENGTH / 2, "NCG_CAR_HALF_WIDTH": K.CAR_WIDTH / 2,
            "NCG_CAR_WHEELBASE": K.CAR_WHEELBASE, "NCG_CAR_MOI": K.CAR_MOMENT_OF_INERTIA, "NCG_CAR_MAX_TORQUE": K.CAR_MAX_TORQUE,
            "NCG_CAR_MAX_POWER": K.CAR_MAX_POWER, "NCG_CAR_MAX_SPEED": K.CAR_MAX_SPEED_MS, "NCG_DRAG_CONSTANT": K.DRAG_CONSTANT,
            "NCG_WEIGHT": K.CAR_MASS * K.GRAVITY, "NCG_STATIC_TYRE_LOAD": K.STATIC_LOAD_PER_TYRE, "NCG_CAR_FRICTION": K.CAR_FRICTION,
            "NCG_CAR_RESTITUTION": K.CAR_RESTITUTION, "NCG_WALL_FRICTION": K.WALL_FRICTION, "NCG_WALL_RESTITUTION": K.WALL_RESTITUTION,
            "NCG_VEL_ITERS": K.VELOCITY_ITERATIONS, "NCG_POS_ITERS": K.POSITION_ITERATIONS, "NCG_STUCK_STEPS": K.STUCK_STEPS,
            "NCG_STUCK_EXT_STEPS": K.STUCK_EXTENDED_STEPS, "NCG_TERMINATE_STEPS": K.TERMINATION_STEPS, "NCG_TRUNCATE_STEPS": K.TRUNCATION_STEPS}
    for k, v in want.items():
        assert float(np.float32(float(defs[k]))) == pytest.approx(float(np.float32(v)), rel=1e-7), k
    assert L.R["NCG_R_USED"] <= L.RECORD_WORDS == 128


def test_spaces_follow_base_env():
    a, o = S.make_spaces(False, 1)
    assert a.shape == (2,) and o.shape == (38,)
    assert a.contains(np.array([0.5, -1.0], dtype=np.float32))
    assert not a.contains(np.array([1.5, 0.0], dtype=np.float32))
    assert not a.contains(np.array([0.5, 0.0], dtype=np.float64))        # gymnasium 0.29: float64 cannot be cast to float32
    a10, o10 = S.make_spaces(False, 10)
    assert a10.shape == (10, 2) and o10.shape == (10, 38)
    d, _ = S.make_spaces(True, 1)
    assert d.contains(4) and not d.contains(5)
    md, _ = S.make_spaces(True, 3)
    assert md.contains(np.array([0, 4, 2])) and not md.contains(np.array([0, 5, 2]))
    assert np.all(o.low[[0, 4, 7, 19, 20, 22]] == np.array([-1, 0, 0, 0, -1, 0], dtype=np.float32))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    """No compute call here (no GPU in the build container): dlopen + symbol lookup only."""
    from nascargymnasium_b200 import engine
    so = engine.build_library()
    lib = ctypes.CDLL(so)
    with open(os.path.join(ROOT, "include", "ncg_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(ncg_[a-z_]+)\s*\(", header.split("typedef struct NcgHandle NcgHandle;")[1]))
    assert declared == set(engine.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ncg_version() == 1


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nascargymnasium_b200.engine import Engine, NcgError
    with pytest.raises((NcgError, ValueError)):
        Engine(4, 1, tracks=["nascar"])


def test_car_env_argument_errors_match_reference():
    from nascargymnasium_b200.car_env import CarEnv
    with pytest.raises(ValueError):
        CarEnv(track_file="tracks/nascar.track", num_cars=11)
    with pytest.raises(ValueError):
        CarEnv(track_file="tracks/nascar.track", num_cars=2, car_names=["a"])
    with pytest.raises(FileNotFoundError):
        CarEnv(track_file="tracks/nope.track")


def test_launch_plan_keeps_ctas_on_one_track_and_fills_the_sms():
    """ncg_plan_ctas (pure host code of the C ABI): whole envs per CTA, <= 32 car slots, never across a track boundary."""
    from nascargymnasium_b200 import engine
    # BASELINE config 2: 4096 single-car envs -> 147 CTAs of <= 28 on 148 SMs (not 128 CTAs of 32)
    first, count = engine.plan_ctas(np.zeros(4096, np.int32), 1, 148)
    assert len(first) == 147 and count.max() == 28 and count.sum() == 4096 and first[0] == 0
    assert np.array_equal(first[1:], np.cumsum(count)[:-1])
    # config 3: 10-car envs -> 3 envs (30 car slots) per CTA
    first, count = engine.plan_ctas(np.zeros(8192, np.int32), 10, 148)
    assert count.max() == 3 and count.sum() == 8192
    # config 4 per GPU: 8192 envs over 8 tracks in blocks -> no CTA straddles a block boundary
    tid = (np.arange(8192) * 8 // 8192).astype(np.int32)
    first, count = engine.plan_ctas(tid, 1, 148)
    assert count.sum() == 8192 and count.max() <= 32
    for f, c in zip(first, count):
        assert len(set(tid[f:f + c].tolist())) == 1
    # worst case (track id alternating per env) degrades to one env per CTA but stays correct
    first, count = engine.plan_ctas((np.arange(64) % 2).astype(np.int32), 1, 148)
    assert len(first) == 64 and (count == 1).all()
    # a single env is one CTA
    first, count = engine.plan_ctas(np.zeros(1, np.int32), 10, 148)
    assert first.tolist() == [0] and count.tolist() == [1]


def test_launch_plan_uses_full_groups_beyond_two_ctas_per_sm():
    """Up to two CTAs per SM a batch is spread over all SMs; beyond, a CTA's step costs about the same with 24 cars as with 32
    (measured in the dispersed steady state, tools/ab_group.sh), so the groups are full and the CTAs as few as possible."""
    from nascargymnasium_b200 import engine
    for E in (8192, 6144):                                     # <= 2 groups of 32 per SM: spread
        first, count = engine.plan_ctas(np.zeros(E, np.int32), 1, 148)
        assert count.sum() == E and count.max() < 32 and len(count) <= 2 * 148
    for E in (12288, 20480, 40960, 65536):
        first, count = engine.plan_ctas(np.zeros(E, np.int32), 1, 148)
        assert count.sum() == E and (count == 32).all() and len(count) == E // 32
    first, count = engine.plan_ctas(np.zeros(8192, np.int32), 10, 148)          # ten-car envs: three envs (30 car slots) per group
    assert count.sum() == 8192 and count.max() == 3 and (count[:-1] == 3).all()


def test_launch_plan_does_not_spill_into_an_extra_wave_on_the_track_mix():
    from nascargymnasium_b200 import engine
    tid = (np.arange(4096) * 8 // 4096).astype(np.int32)         # BASELINE config 4 mix at 4096 envs: 8 blocks of 512
    first, count = engine.plan_ctas(tid, 1, 148)
    assert len(first) <= 148 and count.sum() == 4096 and count.max() <= 32
    for f, c in zip(first, count):
        assert len(set(tid[f:f + c].tolist())) == 1


def test_step_info_builds_final_observation_lazily():
    from nascargymnasium_b200.vector_env import StepInfo
    done = np.array([False, True, False, True])
    rows = np.arange(2 * 38, dtype=np.float32).reshape(2, 38)
    info = StepInfo(done, np.flatnonzero(done), rows, (4, 38), {"r": np.zeros(4), "l": np.zeros(4, int)})
    assert "final_observation" in info and "final_obs" in info and not dict.__contains__(info, "final_observation")
    fo = info["final_observation"]                       # Gymnasium 0.29 form: object array, None for running envs
    assert fo.dtype == object and fo[0] is None and np.array_equal(fo[3], rows[1]) and dict.__contains__(info, "final_observation")
    dense = info["final_obs"]                            # Gymnasium 1.x form
    assert dense.shape == (4, 38) and np.array_equal(dense[1], rows[0]) and not dense[0].any()
    assert info["_final_observation"].tolist() == done.tolist() and info.get("missing") is None
    with pytest.raises(KeyError):
        info["missing"]
