"""Box2D version sensitivity.  box2d-py 2.3.8 bundles some Box2D 2.3.x; the releases differ, for a box
against static boxes, only in b2CollidePolygons (2.3.0: edge-walk b2FindMaxSeparation + k_relativeTol/k_absoluteTol; 2.3.1+:
brute force + k_tol).  Both forms exist in the oracle (oracle/b2lite.h, g_collide_variant) and in the product
(csrc/ncg_b2.cuh, NcgConfig.contacts = 1 | 2).  These tests keep the two implementations of each form in step and pin what
is and is not version-sensitive; tools/b2_version_study.py writes the numbers (profiles/r02_b2_version_study.json)."""
import glob
import os

import numpy as np
import pytest

from nascargymnasium_b200 import track as T
from oracle import oracle as O
from tests import parity_util as P

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traj_*.npz")))


def _replay(g, variant, upto):
    O.set_b2_variant(variant)
    try:
        sp = g["start_pose"]
        env = O.OracleEnv(T.builtin_track_text(str(g["track"])), num_cars=int(g["num_cars"]), reset_on_lap=bool(g["reset_on_lap"]),
                          discrete=bool(g["discrete"]), start_position=(float(sp[0]), float(sp[1])), start_angle=float(sp[2]))
        env.reset()
        out = []
        for t in range(upto):
            a = g["actions"][t]
            o, r, te, tr = env.step(a if not g["discrete"] else a.astype(np.int64))
            out.append(o.copy())
            if g["did_reset"][t]:
                env.reset(fresh=False)
        return np.array(out)
    finally:
        O.set_b2_variant(0)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[5:-4] for p in GOLD])
def test_goldens_are_version_insensitive_up_to_their_first_contact(path):
    """Every golden trajectory is bit-identical under both b2CollidePolygons forms until the car first touches a wall (the
    whole trajectory for the contact-free ones): the fixtures pin the reference's Python half whatever the Box2D minor.
    After the first contacts the two forms may pick different reference faces (2.6 % of touching poses) and the runs part."""
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}
    contact = np.nonzero((g["obs"][:, :, 19] > 0).any(axis=1))[0]
    n = int(contact[0]) if len(contact) else len(g["actions"])
    a, b = _replay(g, 0, n), _replay(g, 1, n)
    assert np.array_equal(a, b)
    assert np.abs(a - g["obs"][:n]).max() < 2e-6


def test_random_poses_disagree_only_in_the_choice_of_reference_face():
    """What differs between the two forms, over random car poses around a track's walls: never whether the boxes touch,
    and whenever they agree on the reference face, point count and features, the coordinates agree to the last bit."""
    env = O.OracleEnv(T.builtin_track_text("martinsville"))
    c = O.b2_variant_study(env, 200000, seed=5)
    assert c["touching"] > 20000
    assert c["touching_differs"] <= 5          # ties at the separation > totalRadius early-out, ~3 per million
    assert c["same_features_bits_differ"] == 0
    assert 0 < c["reference_face_differs"] + c["point_count_differs"] + c["feature_ids_differ"] < 0.05 * c["touching"]


@pytest.mark.parametrize("track,kind,seed", [("martinsville", "drive", 1), ("michigan", "drive", 4)])
def test_device_code_follows_the_230_form_too(track, kind, seed):
    """The host compile of the device code with contacts = 2 against the oracle under variant 1, teacher-forced."""
    recs, act3, raws, exp = P.collect_cases(track, 200, kind=kind, seed=seed, b2_variant=1)
    hc = P.HostCheckEnv(track, contacts=2)
    m = len(recs)
    got = np.zeros_like(recs); obs = np.zeros((m, 38), np.float32); rew = np.zeros(m, np.float32)
    te = np.zeros(m, bool); tr = np.zeros(m, bool)
    for i in range(m):
        hc.records[0] = recs[i]
        o, r, a, b, _ = hc.step(act3[i])
        got[i], obs[i], rew[i], te[i], tr[i] = hc.records[0], o[0], r[0], a, b
    bad, report = P.check_cases(got, obs, rew, te, tr, exp, label=track)
    assert bad == 0, report
    assert (exp["touching"] > 0).sum() > 10


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_engine_matches_the_oracle_under_either_box2d_form(variant):
    from nascargymnasium_b200.engine import Engine
    recs, act3, raws, exp = P.collect_cases("martinsville", 256, kind="drive", seed=11, b2_variant=variant)
    m = len(recs)
    eng = Engine(m, 1, tracks=["martinsville"], auto_reset=False, contacts=1 + variant)
    eng.reset_host()
    eng.set_state_host(recs)
    obs, rew, te, tr, _ = eng.step_host(np.array(raws, dtype=np.float32))
    bad, report = P.check_cases(eng.get_state_host(), obs, rew, te, tr, exp, label=f"variant {variant}")
    assert bad == 0, report
    assert (exp["touching"] > 0).sum() > 20
    eng.close()
