"""Analytic known answers for the optional shared world (car-car contact; oracle/b2lite.h SharedWorld).  The
reference has no such mode (one private b2World per car), so there is nothing of the reference to compare with: these are
Box2D-semantics checks -- restitution of a central hit between equal masses, conservation of linear and angular momentum in an
oblique hit, the listener impulse both cars report -- and the guarantee that the flag changes nothing while it is off."""
import numpy as np
import pytest

from nascargymnasium_b200 import track as T
from oracle import oracle as O

S = O.state_layout()
M, I_Z = 1500.0, 1837.86125


def _env(C=4, **kw):
    return O.OracleEnv(T.builtin_track_text("daytona"), num_cars=C, car_contacts=True, **kw)


def _kin(env, C):
    st = [env.get_state(c) for c in range(C)]
    return np.array([[s[S["S_X"]], s[S["S_Y"]], s[S["S_A"]], s[S["S_VX"]], s[S["S_VY"]], s[S["S_W"]]] for s in st])


def _momentum(k):
    px, py = M * k[:, 3].sum(), M * k[:, 4].sum()
    lz = (M * (k[:, 0] * k[:, 4] - k[:, 1] * k[:, 3]) + I_Z * k[:, 5]).sum()
    return np.array([px, py, lz])


def test_grid_start_and_no_pairs_at_rest():
    env = _env(4)
    k = _kin(env, 4)
    assert np.allclose(k[:, 0], [0.0, 0.0, -8.0, -8.0]) and np.allclose(k[:, 1], [1.5, -1.5, 1.5, -1.5])
    assert env.num_pairs() == 0
    z = np.zeros((4, 2), dtype=np.float32)
    for _ in range(30):
        env.step(z)
    assert env.num_pairs() == 0 and np.allclose(_kin(env, 4)[:, :2], k[:, :2], atol=1e-6)


def test_central_rear_end_hit_between_equal_masses():
    """car 2 (8 m behind car 0, same lane) arrives at 10 m/s on the standing car 0: e = max(0.1, 0.1), so the closing speed
    10 becomes a separating speed 1: v0' = 5.5, v2' = 4.5 (minus a step of drag), and each listener reports m (1 + e) v / 2."""
    env = _env(4)
    s = env.get_state(2); s[S["S_VX"]] = 10.0; env.set_state(s, 2)
    z = np.zeros((4, 2), dtype=np.float32)
    hit = None
    for t in range(60):
        before = _kin(env, 4)
        env.step(z)
        after = _kin(env, 4)
        if env.impulse(0) == 0.0 and after[0, 3] > 1.0 and hit is None:
            hit = t
            break
        if after[0, 3] > 1.0:
            hit = t
            break
    assert hit is not None and 15 < hit < 25                         # 8 m - 5.04 m of car at 10 m/s: 0.3 s
    assert after[0, 3] == pytest.approx(5.5, abs=0.08) and after[2, 3] == pytest.approx(4.5, abs=0.08)
    assert after[0, 3] - after[2, 3] == pytest.approx(0.1 * (before[2, 3] - before[0, 3]), abs=0.02)       # restitution
    assert abs(after[0, 4]) < 1e-3 and abs(after[0, 5]) < 1e-3       # central: no spin, no lateral speed
    assert M * (after[[0, 2], 3].sum() - before[[0, 2], 3].sum()) == pytest.approx(0.0, abs=M * 0.12)     # a step of drag + rolling resistance
    assert np.allclose(after[[1, 3]], before[[1, 3]], atol=1e-4)     # the other lane is not involved


def test_listener_reports_the_pair_impulse_to_both_cars():
    env = _env(4)
    s = env.get_state(2); s[S["S_VX"]] = 10.0; env.set_state(s, 2)
    z = np.zeros((4, 2), dtype=np.float32)
    seen = 0.0
    for t in range(40):
        obs, rew, te, tr = env.step(z)
        # obs[19] = collision impulse / 50000 (of the step), obs[20] = collision angle / pi
        if obs[0, 19] > 0:
            seen = obs[0, 19] * 50000.0
            assert obs[2, 19] == pytest.approx(obs[0, 19], rel=1e-6)
            assert abs(obs[0, 20]) == pytest.approx(1.0, abs=0.02)     # car 0 is hit from behind: the normal points backwards
            assert abs(obs[2, 20]) < 0.02                              # car 2 hits with its nose
            assert rew[0] < -0.5 + 1.0 and rew[2] < 0.5                # the wall-contact penalty applies to car contacts too
            break
    assert seen == pytest.approx(M * 1.1 * 10.0 / 2.0, rel=0.03)


def test_oblique_hit_conserves_linear_and_angular_momentum():
    env = _env(2)
    # car 1 (right lane) steers into car 0: heading 0.5 rad, 8 m/s along its heading
    s = env.get_state(1); s[S["S_A"]] = 0.5; s[S["S_VX"]] = 8.0 * np.cos(0.5); s[S["S_VY"]] = 8.0 * np.sin(0.5); env.set_state(s, 1)
    z = np.zeros((2, 2), dtype=np.float32)
    touched = False
    for t in range(60):
        before = _kin(env, 2)
        env.step(z)
        after = _kin(env, 2)
        if env.num_pairs(touching_only=True):
            touched = True
            d = _momentum(after) - _momentum(before)
            # external forces in one step: drag, rolling resistance, the alignment force (up to ~22 kN for a sliding car) and
            # the damping torque: bounded by |F| dt, far below the ~6000 N s the contact moves between the cars
            assert abs(d[0]) < 1200 and abs(d[1]) < 1200, (t, d)
            exchanged = M * np.abs(after[0, 3:5] - before[0, 3:5]).max()
            if exchanged > 2000:
                assert abs(d[0]) < 0.35 * exchanged and abs(d[1]) < 0.35 * exchanged
    assert touched


def test_flag_off_is_the_reference_semantics():
    a = O.OracleEnv(T.builtin_track_text("daytona"), num_cars=3)
    b = O.OracleEnv(T.builtin_track_text("daytona"), num_cars=3, car_contacts=False)
    rng = np.random.default_rng(0)
    for _ in range(200):
        act = rng.uniform(-1, 1, size=(3, 2)).astype(np.float32)
        oa, ra, _, _ = a.step(act); ob, rb, _, _ = b.step(act)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb)
    assert np.allclose(_kin(a, 3)[:, :2], 0.0, atol=50.0) and a.num_pairs() == 0
