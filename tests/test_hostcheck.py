"""The product's per-car device code (csrc/ncg_car.cuh), compiled for the host by tests/hostcheck, against the oracle.

This is a CPU-side check of the kernel's scalar phases (dynamics, tyres, Box2D step with contacts and TOI, lap timer,
reward, env logic) and of the grid ray traversal; the GPU parity tests in test_gpu_parity.py are the gate proper."""
import ctypes

import numpy as np
import pytest

from nascargymnasium_b200 import layout as L
from nascargymnasium_b200 import track as T
from oracle import oracle as O
from tests import parity_util as P


def _run_cases(track, n, kind, seed, discrete=False):
    recs, act3, raws, exp = P.collect_cases(track, n, kind=kind, seed=seed, discrete=discrete)
    hc = P.HostCheckEnv(track)
    m = len(recs)
    got = np.zeros_like(recs); obs = np.zeros((m, 38), np.float32); rew = np.zeros(m, np.float32)
    te = np.zeros(m, bool); tr = np.zeros(m, bool)
    for i in range(m):
        hc.records[0] = recs[i]
        o, r, a, b, _ = hc.step(act3[i])
        got[i], obs[i], rew[i], te[i], tr[i] = hc.records[0], o[0], r[0], a, b
    return P.check_cases(got, obs, rew, te, tr, exp, label=track), exp


@pytest.mark.parametrize("track,kind,seed", [("nascar", "full", 0), ("martinsville", "drive", 1), ("talladega", "random", 3),
                                             ("michigan", "drive", 4)])
def test_teacher_forced_steps_match_oracle(track, kind, seed):
    (bad, report), exp = _run_cases(track, 250, kind, seed)
    assert bad == 0, report
    if kind != "random":
        assert (exp["touching"] > 0).sum() > 10          # the contact solver / TOI path is exercised


def test_discrete_actions_match_oracle():
    (bad, report), _ = _run_cases("daytona", 150, "random", 5, discrete=True)
    assert bad == 0, report


def test_grid_ray_traversal_matches_box2d_clipping_over_all_walls():
    """The product's grid DDA + slab test against Box2D's own half-plane clipping arithmetic over every wall
    (distance_sensor.py:95-113 ray set-up).  Tolerance 1e-3 m (4e-6 normalised; the stated sensor bound is 1e-3)."""
    hc = P.hostcheck()
    rng = np.random.default_rng(0)
    for name in ("martinsville", "michigan", "trioval", "daytona"):
        tab = T.get_track_table(name)
        blob = np.ascontiguousarray(tab.blob)
        worst = 0.0
        for _ in range(400):
            s = tab.seg64[rng.integers(0, len(tab.seg64))]
            u = rng.uniform()
            x = s[0] + u * (s[2] - s[0]) + rng.uniform(-5, 5)
            y = s[1] + u * (s[3] - s[1]) + rng.uniform(-5, 5)
            th = rng.uniform(-7, 7)
            a, b = np.zeros(16, np.float32), np.zeros(16, np.float32)
            n = ctypes.c_uint(0)
            hc.hc_sensors_brute(P._fp(blob), x, y, th, P._fp(a))
            hc.hc_sensors_grid(P._fp(blob), x, y, th, P._fp(b), ctypes.byref(n))
            d = np.abs(a - b)
            # a ray that grazes a box corner may legitimately hit/miss differently in the two arithmetics
            assert (d > 1e-3).sum() <= 1, (name, x, y, th, a, b)
            worst = max(worst, float(np.sort(d)[-2]))
            assert n.value < 16 * 80
            assert hc.hc_sensors_multi_mismatches(P._fp(blob), x, y, th) == 0
        assert worst < 1e-3, (name, worst)
        # origins outside the grid (every ray scans all walls): the fixed map and the queue must still agree
        for x, y in ((1.0e4, -3.0e3), (float(tab.grid_origin[0]) - 1.0, float(tab.grid_origin[1]) - 1.0)):
            assert hc.hc_sensors_multi_mismatches(P._fp(blob), x, y, 0.3) == 0


def test_rays_in_the_plane_of_a_wall_joint_never_slip_between_the_boxes():
    """Origins exactly in the plane of a box's end face with axis-aligned headings (every car starts on such a pose:
    x = 0, heading 0): some rays then run parallel to the face, in its plane, where the slab arithmetic degenerates.  The
    ray half-length carries a 0.1 mm margin so that two boxes abutting end to end are watertight; Box2D's own clipping
    may let such a ray through (then it reports a farther wall), so the check is one-sided: never farther than Box2D's
    arithmetic by more than 2 mm."""
    hc = P.hostcheck()
    for name in ("daytona", "martinsville", "nascar2"):
        tab = T.get_track_table(name)
        blob = np.ascontiguousarray(tab.blob)
        b = tab.boxes.astype(np.float64)
        for w in range(0, len(b), 3):
            c, s = b[w, 2], b[w, 3]
            ex, ey = b[w, 0] + c * b[w, 4], b[w, 1] + s * b[w, 4]
            for off in (-6.0, 6.0):
                x, y = float(np.float32(ex - s * off)), float(np.float32(ey + c * off))
                for th in (0.0, np.pi / 2, np.pi, -np.pi / 2):
                    a, g = np.zeros(16, np.float32), np.zeros(16, np.float32)
                    n = ctypes.c_uint(0)
                    hc.hc_sensors_brute(P._fp(blob), x, y, float(np.float32(th)), P._fp(a))
                    hc.hc_sensors_grid(P._fp(blob), x, y, float(np.float32(th)), P._fp(g), ctypes.byref(n))
                    assert float((g - a).max()) < 2e-3, (name, w, x, y, th, a, g)


def test_heading_sincos_is_accurate_and_exact_at_zero():
    """csrc/ncg_b2.cuh::sincos_heading (Cody-Waite reduction + minimax polynomials, library fallback for huge angles)
    against float64 sin/cos: <= 2 ulp of 1 over the headings a car can reach and beyond, exact at 0 (the start pose's
    axis-aligned rays rely on it), and still right past the fallback threshold."""
    hc = P.hostcheck()
    hc.hc_sincos_heading.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float)]
    rng = np.random.default_rng(3)
    out = np.zeros(2, np.float32)
    hc.hc_sincos_heading(0.0, P._fp(out))
    assert out[0] == 0.0 and out[1] == 1.0
    worst = 0.0
    for a in np.concatenate([rng.uniform(-8, 8, 4000), rng.uniform(-4.7e4, 4.7e4, 4000), rng.uniform(-1e6, 1e6, 200),
                             np.arange(-40, 41) * (np.pi / 4)]).astype(np.float32):
        hc.hc_sincos_heading(float(a), P._fp(out))
        worst = max(worst, abs(float(out[0]) - np.sin(np.float64(a))), abs(float(out[1]) - np.cos(np.float64(a))))
    assert worst < 2.5e-7, worst


def test_multi_car_env_and_same_track_reset_match_oracle():
    """Free-running 3-car env with reset_on_lap, including reset_car semantics after termination."""
    rng = np.random.default_rng(9)
    orc = O.OracleEnv(T.builtin_track_text("talladega"), num_cars=3, reset_on_lap=True)
    hc = P.HostCheckEnv("talladega", num_cars=3, reset_on_lap=True)
    o0, h0 = orc.reset(), hc.reset()
    assert np.abs(o0 - h0).max() < 1e-6
    resets = 0
    for t in range(900):
        a = np.stack([rng.uniform(0.3, 1.0, 3), rng.uniform(-0.3, 0.7, 3)], axis=1).astype(np.float32)
        # teacher-force every car from the oracle so rounding cannot accumulate
        for c in range(3):
            rec = P.oracle_to_record(orc.get_state(c))
            orc.set_state(P.record_to_oracle(rec), c)
            hc.records[c] = rec
        oo, ro, teo, tro = orc.step(a)
        oh, rh, teh, trh, _ = hc.step(orc.convert_actions(a))
        assert (teo, tro) == (teh, trh), t
        assert np.abs(oo - oh).max() < 1e-3 and np.abs(ro - rh).max() < 1e-3, t
        if teo or tro:
            orc.reset(fresh=False); hc.reset(fresh=False); resets += 1
            for c in range(3):
                want = P.oracle_to_record(orc.get_state(c))
                bad = [b for b in P.compare_records(hc.records[c], want) if b[0] != "NCG_R_LAP_X"]
                assert not bad, bad
    assert resets >= 1


def test_philox_stream_is_the_one_bench_uses_for_the_cpu_baseline():
    import bench
    cars = np.arange(64)
    for step in (0, 1, 999):
        want = bench.synthetic_actions(7, cars, step)
        for c in (0, 5, 63):
            a3 = P.philox_actions_np(7, c, step)
            tb = a3[0] - a3[1]
            assert tb == pytest.approx(want[c, 0], abs=1e-7) and a3[2] == pytest.approx(want[c, 1], abs=1e-7)


def test_nearest_segment_candidate_mask_is_exact():
    """The per-cell candidate mask must never change the answer of the chord-nearest search (banking, progress), on the
    racing surface, in the infield and outside the walls, including the points that fall back to float64."""
    hc = P.hostcheck()
    rng = np.random.default_rng(3)
    for name in T.BUILTIN_TRACK_NAMES:
        tab = T.get_track_table(name)
        blob = np.ascontiguousarray(tab.blob)
        cell = float(blob[T.H_CELL])
        x0, y0 = tab.grid_origin
        nx, ny = tab.grid_dims
        pts = []
        for _ in range(600):                                   # near the centre line
            s = tab.seg64[rng.integers(0, len(tab.seg64))]
            u = rng.uniform()
            pts.append((s[0] + u * (s[2] - s[0]) + rng.uniform(-8, 8), s[1] + u * (s[3] - s[1]) + rng.uniform(-8, 8)))
        for _ in range(600):                                   # anywhere in (and a little outside) the grid
            pts.append((rng.uniform(x0 - 20, x0 + nx * cell + 20), rng.uniform(y0 - 20, y0 + ny * cell + 20)))
        pts += [(0.0, 0.0), (0.0, 1e-6), (1e-3, -1e-6), (-0.5, 0.0)]      # the start-line overlap region
        a, b = np.zeros(2, np.float32), np.zeros(2, np.float32)
        for x, y in pts:
            hc.hc_nearest_segment(P._fp(blob), x, y, 1, P._fp(a))
            hc.hc_nearest_segment(P._fp(blob), x, y, 0, P._fp(b))
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, x, y, a, b)


@pytest.mark.parametrize("track,kind", [("martinsville", "drive"), ("daytona", "drive"), ("michigan", "drive"), ("trioval", "drive")])
def test_toi_early_out_changes_nothing(track, kind):
    """w_solve_toi skips the b2TimeOfImpact query for a broad-phase contact whose start-of-step separation (measured by Collide
    on a separating axis) exceeds what the car can travel in the step.  The claim is that the outcome is the same contact by
    contact: the device source compiled with and without the early-out must produce bit-identical records, observations and
    TOI counters over long free runs with lots of wall contact."""
    rng = np.random.default_rng(3)
    a = P.HostCheckEnv(track)
    b = P.HostCheckEnv(track, no_toi_shortcut=True)
    a.reset(); b.reset()
    touching_steps = skipped_possible = 0
    for t in range(6000):
        act = P.policy_actions(kind, rng, False)
        tb, st = act
        act3 = np.array([max(tb, 0.0), max(-tb, 0.0), st], dtype=np.float32)
        oa, ra, tea, tra, _ = a.step(act3)
        ob, rb, teb, trb, _ = b.step(act3)
        assert np.array_equal(a.records.view(np.uint32), b.records.view(np.uint32)), t
        assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32)) and (tea, tra) == (teb, trb), t
        nc = int(a.records.view(np.uint32)[0, L.R["NCG_R_NCONTACT"]])
        touching_steps += 1 if (nc >> 16) & 0xFFF else 0
        skipped_possible += 1 if (nc & 255) else 0
        if tea or tra:
            a.reset(fresh=False); b.reset(fresh=False)
    assert list(a.counters) == list(b.counters)                   # incl. TOI events and contact steps
    assert skipped_possible > 500, (touching_steps, skipped_possible)


def _rect(cx, cy, a, hx, hy):
    c, s = np.cos(a), np.sin(a)
    l = np.array([[-hx, -hy], [hx, -hy], [hx, hy], [-hx, hy]], dtype=np.float64)
    return np.stack([cx + c * l[:, 0] - s * l[:, 1], cy + s * l[:, 0] + c * l[:, 1]], axis=1)


def _seg_dist(p1, q1, p2, q2):
    """distance of two segments (2-D, float64)"""
    def pt_seg(p, a, b):
        ab = b - a
        t = np.clip(np.dot(p - a, ab) / max(np.dot(ab, ab), 1e-300), 0.0, 1.0)
        return np.linalg.norm(p - (a + t * ab))
    def cross(u, v):
        return u[0] * v[1] - u[1] * v[0]
    d1, d2 = q1 - p1, q2 - p2
    den = cross(d1, d2)
    if abs(den) > 1e-300:
        t = cross(p2 - p1, d2) / den
        u = cross(p2 - p1, d1) / den
        if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0:
            return 0.0
    return min(pt_seg(p1, p2, q2), pt_seg(q1, p2, q2), pt_seg(p2, p1, q1), pt_seg(q2, p1, q1))


def _rect_dist(A, B):
    def inside(p, R):
        s = [(R[(i + 1) % 4][0] - R[i][0]) * (p[1] - R[i][1]) - (R[(i + 1) % 4][1] - R[i][1]) * (p[0] - R[i][0]) for i in range(4)]
        return all(v >= 0 for v in s)
    if inside(A[0], B) or inside(B[0], A):
        return 0.0
    return min(_seg_dist(A[i], A[(i + 1) % 4], B[j], B[(j + 1) % 4]) for i in range(4) for j in range(4))


def test_sweep_face_bound_never_exceeds_the_true_distance_over_the_sweep():
    """The exact TOI early-out rests on sweep_face_bound being a LOWER bound of the distance between the car box and the wall
    box at every time of the sweep (positions and heading interpolated linearly, as b2Sweep does).  Checked here against an
    independent float64 polygon distance sampled along random sweeps: scraping poses, approaches, fast and spinning cars,
    short chords and long straights.  And it is tight where it matters: a car sliding along a wall 1.5 cm away gets ~1.5 cm."""
    lib = P.hostcheck()
    rng = np.random.default_rng(11)
    worst_slack, positive = 1e9, 0
    for case in range(600):
        whx = float(rng.choice([0.7, 2.2, 11.0, 200.0])); why = 0.5
        wang = float(rng.uniform(-np.pi, np.pi)); wpx, wpy = rng.uniform(-800, 800, size=2)
        wc, ws = np.cos(wang), np.sin(wang)
        # the car starts near the wall's inner face, somewhere along it, roughly aligned or at an angle
        along = rng.uniform(-whx - 3, whx + 3); off = why + 1.0 + float(rng.choice([0.016, 0.05, 0.3, 2.0, 6.0])) + rng.uniform(0, 0.02)
        side = rng.choice([-1.0, 1.0])
        c0 = np.array([wpx + wc * along - ws * off * side, wpy + ws * along + wc * off * side])
        a0 = wang + float(rng.choice([0.0, np.pi / 2, rng.uniform(-np.pi, np.pi)])) + rng.uniform(-0.05, 0.05)
        if abs(np.cos(a0 - wang)) < 0.95:
            c0 += np.array([-ws, wc]) * side * 1.6           # not aligned: keep the nose out of the wall
        speed = float(rng.choice([0.0, 1.0, 30.0, 90.0])); hd = rng.uniform(-np.pi, np.pi)
        c1 = c0 + speed / 60.0 * np.array([np.cos(hd), np.sin(hd)])
        a1 = a0 + float(rng.choice([0.0, 1e-3, 0.02, 0.3])) * rng.choice([-1.0, 1.0])
        sw = np.array([c0[0], c0[1], a0, c1[0], c1[1], a1], dtype=np.float32)
        wall = np.array([wpx, wpy, wang, whx, why], dtype=np.float32)
        bound = float(lib.hc_sweep_face_bound(P._fp(sw), P._fp(wall)))
        B = _rect(float(wall[0]), float(wall[1]), float(wall[2]), whx, why)
        dmin = 1e9
        for t in np.linspace(0.0, 1.0, 41):
            c = (1 - t) * sw[0:2].astype(np.float64) + t * sw[3:5].astype(np.float64)
            a = (1 - t) * float(sw[2]) + t * float(sw[5])
            dmin = min(dmin, _rect_dist(_rect(c[0], c[1], a, 2.521, 0.998), B))
        if bound > 0.0:
            positive += 1
            assert bound <= dmin + 2e-4, (case, bound, dmin)          # (float32 rounding at coordinates of ~1e3 m)
            worst_slack = min(worst_slack, dmin - bound)
    assert positive > 150
    # tightness: sliding along a long wall, 1.5 cm between the boxes, no rotation
    sw = np.array([0.0, 0.5 + 0.998 + 0.015, 0.0, 1.3, 0.5 + 0.998 + 0.015, 0.0], dtype=np.float32)
    wall = np.array([0.0, 0.0, 0.0, 200.0, 0.5], dtype=np.float32)
    assert abs(float(lib.hc_sweep_face_bound(P._fp(sw), P._fp(wall))) - 0.015) < 1e-5
