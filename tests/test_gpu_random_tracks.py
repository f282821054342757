"""GPU tests of the reference's random-track mode for a batch (CarEnv(track_file=None): /root/reference/src/car_env.py:264-303,
the way learn/ppo.py:65-77 trains): an env draws a track at reset() and restarts on ANOTHER one whenever an episode ends."""
import numpy as np
import pytest

from nascargymnasium_b200 import layout as L
from nascargymnasium_b200 import track as T

pytestmark = pytest.mark.gpu
R = L.R


def test_finished_envs_move_to_another_track_and_the_histogram_stays_uniform():
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    E, n_tracks = 2048, len(T.BUILTIN_TRACK_NAMES)
    v = NascarVectorEnv(E, track_file=None, discrete_action_space=True)
    obs, _ = v.reset(seed=123)
    t0 = v.track_id.copy()
    assert np.bincount(t0, minlength=n_tracks).min() > 0.6 * E / n_tracks           # every env drew its own track
    v2 = NascarVectorEnv(E, track_file=None, discrete_action_space=True)
    v2.reset(seed=123)
    assert np.array_equal(v2.track_id, t0)                                           # reproducible from the seed
    v2.close()
    # the per-track reset observations, to recognise which track an observation row belongs to
    reset_rows = {}
    for t in range(n_tracks):
        reset_rows[t] = obs[np.flatnonzero(t0 == t)[0]].copy()
        assert all(np.array_equal(obs[e], reset_rows[t]) for e in np.flatnonzero(t0 == t)[:8])
    coast = np.zeros(E, dtype=np.int64)
    for step in range(1, 601):                  # coasting: the stuck rule ends every episode at step 600
        obs, rew, te, tr, info = v.step(coast)
        assert (te.any() or tr.any()) == (step == 600)
    assert te.all()
    t1 = v.track_id
    assert (t1 != t0).all()                                                          # "a different track if possible"
    hist = np.bincount(t1, minlength=n_tracks)
    assert hist.min() > 0.6 * E / n_tracks and hist.max() < 1.5 * E / n_tracks
    for e in range(0, E, 37):                                                        # the row handed out is the new track's reset row
        assert np.array_equal(obs[e], reset_rows[int(t1[e])]), e
    recs = v.engine.get_state_host()
    assert np.array_equal(recs.view(np.uint32)[:, R["NCG_R_TRACK"]], t1.astype(np.uint32))
    assert (recs.view(np.uint32)[:, R["NCG_R_STEP"]] == 0).all()
    # and the next episode runs on the new tracks: compare a few envs with the oracle on their new track
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 5, size=(40, E))
    picks = [int(np.flatnonzero(t1 == t)[0]) for t in range(n_tracks)]
    orcs = {e: O.OracleEnv(T.builtin_track_text(T.BUILTIN_TRACK_NAMES[int(t1[e])]), discrete=True) for e in picks}
    for o in orcs.values():
        o.reset()
    compared = 0
    for k in range(40):
        obs, rew, te, tr, info = v.step(acts[k])
        for e, o in list(orcs.items()):
            oo, ro, teo, tro = o.step([int(acts[k, e])])
            assert (bool(te[e]), bool(tr[e])) == (teo, tro), (k, e)
            assert abs(rew[e] - ro[0]) < 1e-5, (k, e)
            if teo or tro:
                # (random steering at the start line of talladega / michigan / nascar2 / trioval ends an episode within a
                # few steps: the reference's start-line overlap quirk) -- the env has moved on again
                assert np.abs(info["final_obs_rows"][list(info["final_obs_index"]).index(e)] - oo[0]).max() < 1e-4
                del orcs[e]
                continue
            assert np.abs(obs[e] - oo[0]).max() < 1e-4, (k, e)
            compared += 1
    assert compared > 100
    v.close()


def test_redraw_on_the_device_tensor_path_and_staggered_episode_ends():
    """step_torch in random-track mode: envs finish at different steps (half of them coast into the stuck rule at step 600,
    the others drive flat out and end whenever they hit a wall hard enough); the env -> track map follows every time, envs
    that finished are on another track, envs that did not are where they were."""
    import torch
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    E = 600
    v = NascarVectorEnv(E, track_file=None, discrete_action_space=True)
    v.reset_torch(seed=5)
    t0 = v.track_id.copy()
    a = torch.zeros(E, dtype=torch.int32, device="cuda:0")
    a[E // 2:] = 1                                          # full throttle: these envs do not get stuck
    ever = torch.zeros(E, dtype=torch.uint8, device="cuda:0")       # finished at least once (accumulated on the device)
    at600 = None
    for step in range(1, 640):
        obs, rew, te, tr, fin = v.step_torch(a)
        ever |= te | tr
        if step == 600:
            at600 = (te | tr).clone()
    torch.cuda.synchronize()
    t1 = v.track_id
    ever, at600 = ever.cpu().numpy().astype(bool), at600.cpu().numpy().astype(bool)
    coasting = np.arange(E) < E // 2
    assert at600[coasting].all()                             # the stuck rule, all at once
    assert (t1[coasting] != t0[coasting]).all()
    assert (t1[~ever] == t0[~ever]).all()                    # envs that have not finished stay where they were
    once = ever & ~coasting
    assert (t1[once] != t0[once]).mean() > 0.8               # (an env that finished twice may be back on its first track)
    assert torch.equal(v.episode_lengths[:E // 2].cpu(), torch.full((E // 2,), 600, dtype=torch.int32))
    v.close()


def test_fixed_list_and_redraw_off_keep_every_env_on_its_track():
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    E = 256
    for kw in (dict(track_file=["tracks/daytona.track", "tracks/martinsville.track"]), dict(track_file=None, redraw_tracks=False)):
        v = NascarVectorEnv(E, discrete_action_space=True, **kw)
        v.reset(seed=1)
        t0 = v.track_id.copy()
        for step in range(601):
            obs, rew, te, tr, info = v.step(np.zeros(E, dtype=np.int64))
        assert np.array_equal(v.track_id, t0)
        v.close()
