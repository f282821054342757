"""The optional shared world (NcgConfig.car_contacts: the cars of an env collide; SURVEY 8f n3, default off) in the product's
device code against the oracle's SharedWorld (oracle/b2lite.h), free-running from a reset on the start grid: on CPU through the
host compile of the device code (tests/hostcheck), on the GPU through the C ABI.  The analytic known answers of the mode itself
are in tests/test_oracle_carcar.py; here the two implementations have to tell the same story step by step."""
import numpy as np
import pytest

from nascargymnasium_b200 import layout as L
from nascargymnasium_b200 import track as T
from oracle import oracle as O
from tests import parity_util as P

R = L.R


def _actions(kind, C, t, rng):
    a = np.zeros((C, 2), dtype=np.float32)
    if kind == "rear_end":                      # car 2 (second row, left lane) drives into the standing car 0
        a[2] = [1.0, 0.0]
    elif kind == "squeeze":                     # the right lane steers left into the left lane at speed
        a[:, 0] = 0.8
        a[1::2, 1] = 0.35 if t > 60 else 0.0
    else:
        a[:, 0] = rng.uniform(0.3, 1.0, size=C)
        a[:, 1] = rng.uniform(-0.5, 0.5, size=C)
    return a


def _act3(a):
    return np.stack([np.maximum(a[:, 0], 0), np.maximum(-a[:, 0], 0), a[:, 1]], axis=1).astype(np.float32)


@pytest.mark.parametrize("kind,C,steps", [("rear_end", 4, 260), ("squeeze", 4, 200), ("random", 6, 300)])
def test_host_compile_of_the_device_code_follows_the_oracle_shared_world(kind, C, steps):
    rng = np.random.default_rng(1)
    orc = O.OracleEnv(T.builtin_track_text("daytona"), num_cars=C, car_contacts=True)
    hc = P.HostCheckEnv("daytona", num_cars=C, car_contacts=True)
    o0 = hc.reset()
    assert np.abs(o0 - orc.reset()).max() < 1e-6
    touched, worst = 0, 0.0
    for t in range(steps):
        a = _actions(kind, C, t, rng)
        oo, ro, teo, tro = orc.step(a)
        oh, rh, teh, trh, _ = hc.step(_act3(a))
        npairs_t = orc.num_pairs(touching_only=True)
        touched += 1 if npairs_t else 0
        worst = max(worst, float(np.abs(oh - oo).max()))
        # free-running float32 against float32: identical operation order, FMA contraction off on both sides
        assert np.abs(oh - oo).max() < 2e-4, (t, int(np.abs(oh - oo).argmax()), worst)
        assert np.abs(rh - ro).max() < 1e-3 and (teh, trh) == (teo, tro), t
        n_exist = int(hc.pairs.view(np.uint32)[8 * 45])
        assert n_exist == orc.num_pairs(), t
        if teo or tro:
            break
    assert touched > 5, (kind, touched)


def test_reset_car_in_the_middle_of_a_car_contact():
    """CarPhysics.reset_car semantics in the shared world: the cars are put back on the grid while two of them touch; the stale
    pair ends (EndContact on both listeners) when the next step finds the fat AABBs apart, on both sides alike."""
    rng = np.random.default_rng(1)
    C = 4
    orc = O.OracleEnv(T.builtin_track_text("daytona"), num_cars=C, car_contacts=True)
    hc = P.HostCheckEnv("daytona", num_cars=C, car_contacts=True)
    hc.reset(); orc.reset()
    touched = 0
    for t in range(200):
        a = _actions("rear_end", C, t, rng)
        oo, _, _, _ = orc.step(a)
        oh, _, _, _, _ = hc.step(_act3(a))
        touched += 1 if orc.num_pairs(touching_only=True) else 0
        if touched >= 3:
            break
    assert touched >= 3 and orc.num_pairs(touching_only=True)
    assert np.abs(hc.reset(fresh=False) - orc.reset(fresh=False)).max() < 1e-6
    for t in range(150):
        a = _actions("rear_end", C, t, rng)
        oo, ro, teo, tro = orc.step(a)
        oh, rh, teh, trh, _ = hc.step(_act3(a))
        assert np.abs(oh - oo).max() < 2e-4 and np.abs(rh - ro).max() < 1e-3, t
        assert int(hc.pairs.view(np.uint32)[8 * 45]) == orc.num_pairs(), t
        if t == 0:
            assert orc.num_pairs(touching_only=True) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,C,steps", [("rear_end", 4, 220), ("squeeze", 4, 200), ("random", 6, 300)])
def test_engine_follows_the_oracle_shared_world(kind, C, steps):
    from nascargymnasium_b200.engine import Engine
    rng = np.random.default_rng(1)
    E = 40                                                     # the same env 40 times over: every copy must agree
    orc = O.OracleEnv(T.builtin_track_text("daytona"), num_cars=C, car_contacts=True)
    eng = Engine(E, C, tracks=["daytona"], auto_reset=False, car_contacts=True)
    o0 = eng.reset_host().reshape(E, C, 38)
    assert np.abs(o0 - orc.reset()[None]).max() < 1e-6
    touched = 0
    for t in range(steps):
        a = _actions(kind, C, t, rng)
        oo, ro, teo, tro = orc.step(a)
        og, rg, teg, trg, _ = eng.step_host(np.broadcast_to(a, (E, C, 2)).copy())
        og = og.reshape(E, C, 38)
        assert np.array_equal(og[0].view(np.uint32), og[-1].view(np.uint32)) and np.array_equal(og[0].view(np.uint32), og[E // 2].view(np.uint32))
        touched += 1 if orc.num_pairs(touching_only=True) else 0
        # the engine contracts FMAs, the oracle does not: a contact amplifies the last-bit differences
        assert np.abs(og[0] - oo).max() < (2e-4 if not touched else 5e-3), (t, touched, int(np.abs(og[0] - oo).argmax()))
        assert (bool(teg[0]), bool(trg[0])) == (teo, tro), t
        if teo or tro:
            break
    assert touched > 5
    eng.close()


@pytest.mark.gpu
def test_car_contacts_off_is_bit_identical_and_on_conserves_momentum():
    import torch
    from nascargymnasium_b200.engine import Engine
    E, C = 64, 10
    outs = []
    for _ in range(2):
        eng = Engine(E, C, tracks=["talladega"], auto_reset=True)
        eng.reset_host()
        o = torch.empty((E * C, 38), device="cuda:0")
        eng.rollout(300, seed=3, mode=1, obs_last=o.view(-1))
        torch.cuda.synchronize()
        outs.append(eng.get_state_host().copy())
        eng.close()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    # flag on: ten cars driving off a 2 x 5 grid bump into each other; nothing blows up, cars stay apart
    eng = Engine(E, C, tracks=["talladega"], auto_reset=True, car_contacts=True)
    eng.reset_host()
    o = torch.empty((E * C, 38), device="cuda:0")
    eng.rollout(400, seed=3, mode=1, obs_last=o.view(-1))
    torch.cuda.synchronize()
    rec = eng.get_state_host().reshape(E, C, -1)
    # (float words only: a listener entry of a car-car contact holds the other car as the integer -1 - j)
    fw = [R[k] for k in ("NCG_R_X", "NCG_R_Y", "NCG_R_ANGLE", "NCG_R_VX", "NCG_R_VY", "NCG_R_OMEGA", "NCG_R_RPM", "NCG_R_SLIP",
                         "NCG_R_CUM_IMPACT", "NCG_R_CUM_REWARD", "NCG_R_IMPULSE")] + [R["NCG_R_TYRE_TEMP"] + i for i in range(12)]
    assert np.isfinite(rec[:, :, fw]).all() and torch.isfinite(o).all()
    xy = rec[:, :, [R["NCG_R_X"], R["NCG_R_Y"]]].astype(np.float64)
    d = np.linalg.norm(xy[:, :, None, :] - xy[:, None, :, :], axis=-1) + np.eye(C)[None] * 1e9
    assert d.min() > 1.9                                      # two car boxes (5.04 x 2.0 m) cannot be closer than a car width
    assert not np.array_equal(rec.reshape(E * C, -1).view(np.uint32), outs[0].view(np.uint32))
    eng.close()


def _kin(rec, C):
    return rec.reshape(-1, C, rec.shape[-1])[0][:, [R["NCG_R_X"], R["NCG_R_Y"], R["NCG_R_ANGLE"], R["NCG_R_VX"], R["NCG_R_VY"], R["NCG_R_OMEGA"]]].astype(np.float64)


@pytest.mark.gpu
def test_engine_known_answers_of_the_shared_world():
    """The analytic cases of tests/test_oracle_carcar.py on the CUDA engine itself: a central rear-end hit between equal masses
    (e = 0.1: closing 10 m/s -> separating 1 m/s, 5.5 / 4.5 m/s), both listeners report m (1 + e) v / 2, and a chain of three
    (car 4 hits car 2, which then hits car 0) moves momentum forward without creating any."""
    from nascargymnasium_b200.engine import Engine
    M = 1500.0
    C = 6
    eng = Engine(2, C, tracks=["daytona"], auto_reset=False, car_contacts=True)
    eng.reset_host()
    rec = eng.get_state_host().reshape(2, C, -1)
    rec[0, 2, R["NCG_R_VX"]] = 10.0                             # env 0: car 2 runs into the standing car 0
    rec[1, 4, R["NCG_R_VX"]] = 14.0                             # env 1: car 4 -> car 2 -> car 0
    eng.set_state_host(rec.reshape(2 * C, -1))
    z = np.zeros((2, C, 2), dtype=np.float32)
    hit, imp, moved0 = None, 0.0, None
    for t in range(120):
        before = eng.get_state_host().reshape(2, C, -1)
        obs, rew, te, tr, _ = eng.step_host(z)
        after = eng.get_state_host().reshape(2, C, -1)
        obs = obs.reshape(2, C, 38)
        if hit is None and after[0, 0, R["NCG_R_VX"]] > 1.0:
            hit = t
            kb, ka = _kin(before[0:1], C), _kin(after[0:1], C)
            assert ka[0, 3] == pytest.approx(5.5, abs=0.08) and ka[2, 3] == pytest.approx(4.5, abs=0.08)
            assert ka[0, 3] - ka[2, 3] == pytest.approx(0.1 * (kb[2, 3] - kb[0, 3]), abs=0.02)
            assert abs(ka[0, 4]) < 1e-3 and abs(ka[0, 5]) < 1e-3
            assert np.allclose(ka[[1, 3, 5]], kb[[1, 3, 5]], atol=1e-4)
            imp = obs[0, 0, 19] * 50000.0
            assert obs[0, 2, 19] == pytest.approx(obs[0, 0, 19], rel=1e-6)
            assert abs(obs[0, 0, 20]) == pytest.approx(1.0, abs=0.02) and abs(obs[0, 2, 20]) < 0.02
        if moved0 is None and after[1, 0, R["NCG_R_VX"]] > 1.0:
            moved0 = t
        px = M * after[1, :, R["NCG_R_VX"]].astype(np.float64).sum()
        assert px < M * 14.0 + 1.0                              # momentum is handed on, never created
    assert hit is not None and 15 < hit < 25
    assert imp == pytest.approx(M * 1.1 * 10.0 / 2.0, rel=0.03)
    assert moved0 is not None and moved0 > hit                  # the second link of the chain is hit later
    last = eng.get_state_host().reshape(2, C, -1)[1]
    assert last[0, R["NCG_R_X"]] > 0.5 and last[2, R["NCG_R_X"]] > -8.0 + 0.5 and last[0, R["NCG_R_X"]] - last[2, R["NCG_R_X"]] > 5.0
    eng.close()


@pytest.mark.gpu
def test_checkpoint_of_a_shared_world_engine_in_the_middle_of_a_contact():
    """get_state_host + get_car_pairs_host is the whole state: an engine restored from them in the middle of a car-car contact
    continues bit for bit like the one that kept running."""
    from nascargymnasium_b200.engine import Engine
    rng = np.random.default_rng(1)
    E, C = 8, 4
    a_eng = Engine(E, C, tracks=["daytona"], auto_reset=False, car_contacts=True)
    b_eng = Engine(E, C, tracks=["daytona"], auto_reset=False, car_contacts=True)
    a_eng.reset_host(); b_eng.reset_host()
    touched = 0
    for t in range(220):
        a = np.broadcast_to(_actions("rear_end", C, t, rng), (E, C, 2)).copy()
        a_eng.step_host(a)
        pairs = a_eng.get_car_pairs_host()
        touched += 1 if (pairs.view(np.uint32)[0, 0::8][:45] & 2).any() else 0
        if touched == 4:
            break
    assert touched == 4
    b_eng.set_state_host(a_eng.get_state_host()); b_eng.set_car_pairs_host(pairs)
    for t in range(60):
        a = np.broadcast_to(_actions("rear_end", C, t, rng), (E, C, 2)).copy()
        oa = a_eng.step_host(a)[0]; ob = b_eng.step_host(a)[0]
        assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32)), t
    assert np.array_equal(a_eng.get_state_host().view(np.uint32), b_eng.get_state_host().view(np.uint32))
    assert np.array_equal(a_eng.get_car_pairs_host().view(np.uint32), b_eng.get_car_pairs_host().view(np.uint32))
    a_eng.close(); b_eng.close()
