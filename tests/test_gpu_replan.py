"""GPU tests of the host-side launch plan following the env -> track map on a live engine (the CTA and pair
tables have separate capacities; set_state / reset validate track ids before anything is launched)."""
import numpy as np
import pytest

from nascargymnasium_b200 import layout as L
from nascargymnasium_b200 import track as T

pytestmark = pytest.mark.gpu
R = L.R


def _rollout_obs(eng, steps, seed):
    import torch
    o = torch.empty((eng.num_cars, 38), dtype=torch.float32, device=f"cuda:{eng.device}")
    eng.rollout(steps, seed=seed, obs_last=o)
    torch.cuda.synchronize()
    return o.cpu().numpy()


def test_replanning_a_live_engine_from_one_track_to_eight_and_back():
    from nascargymnasium_b200.engine import Engine
    E = 4096
    names = list(T.BUILTIN_TRACK_NAMES)
    eng = Engine(E, 1, tracks=names)
    one = np.zeros(E, dtype=np.int32)
    mix = (np.arange(E, dtype=np.int64) * len(names) // E).astype(np.int32)
    inter = (np.arange(E) % len(names)).astype(np.int32)            # worst case: every env its own CTA boundary
    eng.reset_host(track_id=one)
    a = _rollout_obs(eng, 5, 1)
    for tid in (mix, inter, one, mix):
        eng.reset_host(track_id=tid)                               # grows / shrinks both tables on the live handle
        o = _rollout_obs(eng, 5, 1)
        assert np.isfinite(o).all()
        recs = eng.get_state_host()
        assert (recs.view(np.uint32)[:, R["NCG_R_TRACK"]] == tid.astype(np.uint32)).all()
    eng.reset_host(track_id=one)
    assert np.isfinite(_rollout_obs(eng, 5, 1)).all() and np.isfinite(a).all()
    eng.close()


def test_interleaved_track_map_equals_blocked_map_env_by_env():
    """an env's trajectory does not depend on which CTA serves it: interleaved ids vs the same envs sorted by track"""
    from nascargymnasium_b200.engine import Engine
    import torch
    E = 512
    names = list(T.BUILTIN_TRACK_NAMES)
    inter = (np.arange(E) % len(names)).astype(np.int32)
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, size=(20, E, 2)).astype(np.float32)
    e1 = Engine(E, 1, tracks=names)
    e1.reset_host(track_id=inter)
    order = np.argsort(inter, kind="stable")
    e2 = Engine(E, 1, tracks=names)
    e2.reset_host(track_id=inter[order])
    for t in range(20):
        o1 = e1.step_host(acts[t])[0]
        o2 = e2.step_host(acts[t][order])[0]
        assert np.array_equal(o1[order], o2)
    e1.close(); e2.close()


def test_set_state_on_device_follows_and_validates_the_track_map():
    from nascargymnasium_b200.engine import Engine
    import torch
    E = 600
    names = list(T.BUILTIN_TRACK_NAMES)
    a = Engine(E, 1, tracks=names)
    a.reset_host(track_id=(np.arange(E) % len(names)).astype(np.int32))
    rng = np.random.default_rng(1)
    acts = rng.uniform(-1, 1, size=(8, E, 2)).astype(np.float32)
    for t in range(4):
        a.step_host(acts[t])
    st = a.get_state()
    b = Engine(E, 1, tracks=names)
    b.reset_host(track_id=np.zeros(E, dtype=np.int32))              # a different map: b must re-plan from the records
    b.set_state(st)
    for t in range(4, 8):
        oa = a.step_host(acts[t])[0]
        ob = b.step_host(acts[t])[0]
        assert np.array_equal(oa, ob)
    bad = st.clone()
    bad.view(torch.int32)[7, R["NCG_R_TRACK"]] = 99
    before = b.get_state_host()
    with pytest.raises(ValueError):
        b.set_state(bad)
    assert np.array_equal(before.view(np.uint32), b.get_state_host().view(np.uint32))       # nothing was written
    with pytest.raises(ValueError):
        b.get_state(out=torch.empty(5, device="cuda:0"))
    # device-side reset ids are checked before the kernel follows them
    tid = torch.full((E,), 3, dtype=torch.int32, device="cuda:0"); tid[11] = 8
    with pytest.raises(ValueError):
        b.reset(track_id=tid, fresh=True)
    assert np.array_equal(before.view(np.uint32), b.get_state_host().view(np.uint32))
    a.close(); b.close()


def test_vector_env_rejects_invalid_actions_like_the_reference():
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    v = NascarVectorEnv(16, track_file="tracks/daytona.track")
    v.reset()
    with pytest.raises(AssertionError):
        v.step(np.full((16, 2), 1.5, dtype=np.float32))
    with pytest.raises(AssertionError):
        v.step(np.zeros((16, 2), dtype=np.float64))
    a = np.zeros((16, 2), dtype=np.float32); a[3, 0] = np.nan
    with pytest.raises(AssertionError):
        v.step(a)
    v.step(np.zeros((16, 2), dtype=np.float32))
    v.close()
    d = NascarVectorEnv(16, track_file="tracks/daytona.track", discrete_action_space=True)
    d.reset()
    with pytest.raises(AssertionError):
        d.step(np.full(16, 5, dtype=np.int64))
    d.step(np.full(16, 4, dtype=np.int64))
    d.close()


def test_rank_slices_with_a_car_base_reproduce_the_single_engine_job():
    """bench.py shards a job with distributed.shard_range and gives every rank the Philox streams of its own global car
    indices (ncg_set_rollout_base): the union of the slices is bit-identical to one engine running the whole job."""
    import torch
    from nascargymnasium_b200 import distributed as D
    from nascargymnasium_b200.engine import Engine
    E, C, W, steps = 1200, 2, 3, 40
    names = list(T.BUILTIN_TRACK_NAMES)
    full = Engine(E, C, tracks=names)
    full.reset_host(track_id=np.sort(np.arange(E) % len(names)).astype(np.int32))
    o = torch.empty((E * C, 38), dtype=torch.float32, device="cuda:0")
    full.rollout(steps, seed=7, obs_last=o)
    ref = o.cpu().numpy().reshape(E, C, 38)
    ref_tid = full.get_state_host().view(np.uint32).reshape(E, C, -1)[:, 0, R["NCG_R_TRACK"]]
    got = {}
    for r in range(W):
        lo, hi = D.shard_range(E, r, W)
        tid = D.shard_track_ids(E, len(names), r, W)
        eng = Engine(hi - lo, C, tracks=names)
        eng.reset_host(track_id=tid)
        eng.set_rollout_base(car_base=lo * C)
        oo = torch.empty(((hi - lo) * C, 38), dtype=torch.float32, device="cuda:0")
        eng.rollout(steps, seed=7, obs_last=oo)
        got[r] = (tid, oo.cpu().numpy().reshape(hi - lo, C, 38))
        eng.close()
    # a slice's env j is global car stream (lo + j) on track tid[j]: compare with the single engine stepping the same
    # (stream, track) pairs -- build that engine with the concatenated map
    cat_tid = np.concatenate([got[r][0] for r in range(W)])
    both = Engine(E, C, tracks=names)
    both.reset_host(track_id=cat_tid)
    both.rollout(steps, seed=7, obs_last=o)
    whole = o.cpu().numpy().reshape(E, C, 38)
    assert np.array_equal(whole, np.concatenate([got[r][1] for r in range(W)]))
    assert np.isfinite(ref).all() and ref_tid.max() == len(names) - 1
    full.close(); both.close()


def test_step_torch_reports_episode_return_and_length_on_the_device():
    import torch
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    E = 64
    v = NascarVectorEnv(E, track_file="tracks/martinsville.track", discrete_action_space=True)
    v.reset_torch()
    a = torch.zeros(E, dtype=torch.int32, device="cuda:0")            # coasting: stuck rule ends every episode at step 600
    ret = torch.zeros(E, device="cuda:0")
    for t in range(1, 601):
        obs, rew, te, tr, fin = v.step_torch(a)
        ret += rew
    torch.cuda.synchronize()
    assert bool(te.all()) and not bool(tr.any())
    assert torch.equal(v.episode_lengths.cpu(), torch.full((E,), 600, dtype=torch.int32))
    assert torch.allclose(v.episode_returns.cpu(), ret.cpu(), atol=1e-3)
    assert float(v.episode_returns[0]) == pytest.approx(-0.05 * 599 + 10.0 - 0.0, abs=0.1)
    v.close()
