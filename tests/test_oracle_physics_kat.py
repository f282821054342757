"""Known-answer checks of the oracle's rigid-body half (oracle/b2lite.h) against what Box2D's published algorithm must
produce in analytically simple situations.  The reference ships no tests and box2d-py is not installable offline, so
the Box2D half stays "parity unpinned" (DESIGN.md section 4); these cases pin the parts of it that have a closed form:
restitution mixing and the velocity threshold, the summed normal impulse the collision listener reports, and that the
TOI sub-stepping keeps a fast car out of a 1 m thick wall.  Box2D facts used (b2Settings.h / b2Contact.h):
restitution = max(0.1, 0.25), b2_velocityThreshold = 1 m/s, b2_linearSlop = 0.005, b2_maxTranslation = 2 m per step."""
import math

import numpy as np
import pytest

from nascargymnasium_b200 import constants as K
from nascargymnasium_b200 import track as T
from oracle import oracle as O

S = O.state_layout()
HALF_LEN, HALF_WID = 5.042 / 2, 1.996 / 2
WALL_FACE_Y = 5.5          # daytona's first straight: wall boxes centred on y = +-6, 0.5 m half thickness


def _place(env, x, y, angle, vx, vy):
    """Teleport the freshly reset car: pose, velocity and a proxy AABB around the new pose that is flagged as moved."""
    s = env.get_state().copy()
    s[S["S_X"]], s[S["S_Y"]], s[S["S_A"]], s[S["S_VX"]], s[S["S_VY"]], s[S["S_W"]] = x, y, angle, vx, vy, 0.0
    ext = abs(math.cos(angle)) * HALF_LEN + abs(math.sin(angle)) * HALF_WID + 0.11
    eyt = abs(math.sin(angle)) * HALF_LEN + abs(math.cos(angle)) * HALF_WID + 0.11
    s[S["S_FLX"]], s[S["S_FLY"]], s[S["S_FUX"]], s[S["S_FUY"]] = x - ext, y - eyt, x + ext, y + eyt
    s[S["S_PROXYMOVED"]] = 1
    s[S["S_NCONTACT"]] = 0
    s[S["S_PVX"]], s[S["S_PVY"]] = vx, vy
    s[S["S_PREVX"]], s[S["S_PREVY"]], s[S["S_LX"]], s[S["S_LY"]] = x, y, x, y
    env.set_state(s)


def _run_until_bounce(v0, max_steps=200):
    env = O.OracleEnv(T.builtin_track_text("daytona"))
    env.reset()
    _place(env, 50.0, 0.0, math.pi / 2, 0.0, v0)          # nose towards the wall at y = +5.5, 2.98 m of free travel
    hist = []
    for _ in range(max_steps):
        obs, _, _, _ = env.step([[0.0, 0.0]])
        s = env.get_state()
        # obs[19] = collision impulse / 50000 (the listener's value is cleared at the end of the step)
        hist.append((s[S["S_Y"]], s[S["S_VY"]], s[S["S_VX"]], s[S["S_W"]], float(obs[0][19]) * 50000.0))
        if s[S["S_VY"]] < 0:
            break
    return np.array(hist)


@pytest.mark.parametrize("v0", [5.0, 20.0])
def test_head_on_bounce_has_restitution_one_quarter_and_reports_the_summed_impulse(v0):
    h = _run_until_bounce(v0)
    assert h[-1, 1] < 0, "the car never bounced"
    v_in = h[-2, 1] if len(h) > 1 else v0                   # speed going into the contact step (drag has taken a little)
    v_out = -h[-1, 1]
    assert v_out == pytest.approx(0.25 * v_in, rel=0.03)    # b2MixRestitution(0.1, 0.25) = 0.25
    assert abs(h[-1, 2]) < 1e-2 and abs(h[-1, 3]) < 1e-2     # symmetric two-point manifold: no sideways or angular kick (float32 angle = pi/2)
    # CarCollisionListener.PostSolve sums the normal impulses of the manifold points: m (1 + e) v
    assert h[-1, 4] == pytest.approx(K.CAR_MASS * 1.25 * v_in, rel=0.03)
    # and the nose never got past the wall face by more than the solver's slop
    assert h[:, 0].max() + HALF_LEN < WALL_FACE_Y + 0.03


def test_slow_contact_is_inelastic_below_the_velocity_threshold():
    """|v_n| < b2_velocityThreshold: the solver adds no restitution bias, so a 0.8 m/s touch just stops the car."""
    env = O.OracleEnv(T.builtin_track_text("daytona"))
    env.reset()
    _place(env, 50.0, WALL_FACE_Y - HALF_LEN - 0.2, math.pi / 2, 0.0, 0.8)      # nose 0.2 m from the wall face
    vy = []
    for _ in range(60):
        env.step([[0.0, 0.0]])
        s = env.get_state()
        vy.append(s[S["S_VY"]])
        assert s[S["S_Y"]] + HALF_LEN < WALL_FACE_Y + 0.03
    assert env.num_contacts()[1] >= 1                      # it is touching the wall
    assert min(vy) > -0.05 and abs(vy[-1]) < 0.02          # no bounce, at rest


def test_time_of_impact_keeps_a_fast_car_out_of_the_wall():
    """110 m/s = 1.83 m per step against a 1 m thick wall: without TOI sub-stepping the nose would end a step inside or
    beyond the box; b2World::SolveTOI has to stop it at the face."""
    env = O.OracleEnv(T.builtin_track_text("daytona"))
    env.reset()
    _place(env, 50.0, 0.0, math.pi / 2, 0.0, 110.0)
    worst = -1e9
    for _ in range(12):
        env.step([[0.0, 0.0]])
        s = env.get_state()
        worst = max(worst, s[S["S_Y"]] + HALF_LEN - WALL_FACE_Y)
    assert worst < 0.05, worst
    assert s[S["S_VY"]] < 0                                  # it bounced back
