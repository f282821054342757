"""GPU tests of the drop-in surfaces (CarEnv, NascarVectorEnv, rollout kernel, auto-reset) through the C ABI."""
import numpy as np
import pytest

from nascargymnasium_b200 import layout as L
from nascargymnasium_b200 import track as T
from tests import parity_util as P

pytestmark = pytest.mark.gpu
R, F = L.R, L.F


def test_car_env_mirror_matches_oracle_and_reference_info_keys():
    from nascargymnasium_b200.car_env import CarEnv
    from oracle import oracle as O
    env = CarEnv(track_file="tracks/nascar.track")
    with pytest.raises(RuntimeError):
        env.step(np.zeros(2, dtype=np.float32))
    obs, info = env.reset(seed=3)
    orc = O.OracleEnv(T.builtin_track_text("nascar"))
    assert np.abs(obs - orc.reset()[0]).max() < 1e-6 and obs.shape == (38,) and obs.dtype == np.float32
    for k in ("simulation_time", "num_cars", "followed_car_index", "termination_reason", "cars", "physics"):
        assert k in info
    for k in ("car_index", "disabled", "car_position", "car_speed_kmh", "car_speed_ms", "on_track", "performance", "lap_timing",
              "cumulative_reward", "cumulative_impact_force"):
        assert k in info["cars"][0]
    assert set(info["cars"][0]["lap_timing"]) == {"current_lap_time", "last_lap_time", "best_lap_time", "lap_count", "is_timing",
                                                  "has_crossed_startline", "total_distance_traveled", "formatted_current",
                                                  "formatted_last", "formatted_best"}
    with pytest.raises(AssertionError):
        env.step(np.array([2.0, 0.0], dtype=np.float32))
    rng = np.random.default_rng(0)
    ret = 0.0
    for t in range(300):
        a = np.array([rng.uniform(0.3, 1), rng.uniform(-0.1, 0.1)], dtype=np.float32)
        o, r, te, tr, info = env.step(a)
        oo, ro, teo, tro = orc.step([a])
        assert isinstance(te, bool) and isinstance(tr, bool) and o.shape == (38,)
        assert np.abs(o - oo[0]).max() < 2e-3 and abs(r - ro[0]) < 1e-3 and (te, tr) == (teo, tro)
        assert info["cars"][0]["on_track"] == orc.on_track()
        ret += float(r)
    assert info["simulation_time"] == pytest.approx(orc.sim_time, abs=1e-12)
    assert info["cars"][0]["cumulative_reward"] == pytest.approx(ret, abs=1e-2)
    # same-track reset goes through reset_car semantics and returns the start observation
    obs2, _ = env.reset()
    assert np.abs(obs2 - obs).max() < 1e-6
    env.close()


def test_stuck_rule_disables_at_step_600_with_plus_ten():
    """zero action -> -0.05 per step, +10 and terminated at step 600 (all_cars_disabled)."""
    from nascargymnasium_b200.car_env import CarEnv
    env = CarEnv(track_file="tracks/daytona.track", discrete_action_space=True)
    env.reset()
    total = 0.0
    for t in range(1, 601):
        o, r, te, tr, info = env.step(0)
        total += float(r)
        if t < 600:
            assert not te and r == pytest.approx(-0.05)
    assert te and not tr and r == pytest.approx(10.0)
    assert info["termination_reason"] == "all_cars_disabled" and env.disabled_cars == {0}
    assert total == pytest.approx(10 - 599 * 0.05, abs=1e-3)
    env.close()


def test_multi_car_env_matches_oracle():
    from nascargymnasium_b200.car_env import CarEnv
    from oracle import oracle as O
    C = 10
    env = CarEnv(track_file="tracks/talladega.track", num_cars=C)
    orc = O.OracleEnv(T.builtin_track_text("talladega"), num_cars=C)
    obs, _ = env.reset()
    assert obs.shape == (C, 38) and np.abs(obs - orc.reset()).max() < 1e-6
    rng = np.random.default_rng(4)
    for t in range(250):
        a = np.stack([rng.uniform(0.2, 1.0, C), rng.uniform(-0.2, 0.6, C)], axis=1).astype(np.float32)
        o, r, te, tr, info = env.step(a)
        oo, ro, teo, tro = orc.step(a)
        assert (te, tr) == (teo, tro), t
        assert r.shape == (C,) and np.abs(r - ro).max() < 2e-3, t
        assert np.abs(o - oo).max() < 5e-3, t
        if te or tr:
            env.reset(); orc.reset(fresh=False)
    env.close()


def test_vector_env_autoreset_and_final_observation():
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    E = 64
    venv = NascarVectorEnv(E, track_file="tracks/martinsville.track", discrete_action_space=True)
    obs0, _ = venv.reset()
    assert obs0.shape == (E, 38)
    acts = np.zeros(E, dtype=np.int64)
    acts[::2] = 1                                       # odd envs never move -> stuck-disabled at step 600
    seen_done = False
    for t in range(1, 602):
        obs, rew, te, tr, info = venv.step(acts)
        if t == 600:
            assert te[1::2].all() and not te[::2].any()
            assert info["_final_observation"].tolist() == te.tolist()
            assert np.allclose(rew[1::2], 10.0)
            # finished envs come back already reset: observation equals the start observation, final obs does not
            assert np.abs(obs[1::2] - obs0[1::2]).max() < 1e-6
            assert info["episode"]["l"][1] == 600 and info["episode"]["r"][1] == pytest.approx(10 - 599 * 0.05, abs=1e-3)
            seen_done = True
    assert seen_done
    venv.close()


def test_rollout_kernel_equals_single_steps_with_the_same_philox_actions():
    import bench
    import torch
    from nascargymnasium_b200.engine import Engine
    E, Tn = 256, 40
    a = Engine(E, 1, tracks=["michigan"], auto_reset=True)
    b = Engine(E, 1, tracks=["michigan"], auto_reset=True)
    a.reset_host(); b.reset_host()
    obs_roll = torch.empty((Tn, E, 38), dtype=torch.float32, device="cuda:0")
    rew_roll = torch.empty((Tn, E), dtype=torch.float32, device="cuda:0")
    a.rollout(Tn, seed=5, mode=0, obs_rollout=obs_roll.view(-1), reward_rollout=rew_roll.view(-1))
    torch.cuda.synchronize()
    cars = np.arange(E)
    for t in range(Tn):
        obs, rew, te, tr, _ = b.step_host(bench.synthetic_actions(5, cars, t))
        assert np.abs(obs - obs_roll[t].cpu().numpy()).max() < 1e-6, t
        assert np.abs(rew - rew_roll[t].cpu().numpy()).max() < 1e-6, t
    assert np.array_equal(a.get_state_host(), b.get_state_host())
    a.close(); b.close()


def test_torch_device_path_keeps_observations_on_device():
    import torch
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    venv = NascarVectorEnv(512, track_file=None)        # all 8 tracks, sorted by track id
    obs = venv.reset_torch()
    assert obs.is_cuda and obs.shape == (512, 38)
    act = torch.rand((512, 2), device="cuda:0") * 2 - 1
    for _ in range(20):
        obs, rew, te, tr, fin = venv.step_torch(act)
    torch.cuda.synchronize()
    assert obs.is_cuda and torch.isfinite(obs).all() and rew.shape == (512,)
    st = venv.engine.read_stats()
    assert st["car_steps"] == 512 * 20
    recs = venv.engine.get_state_host()
    assert sorted(set(recs.view(np.uint32)[:, R["NCG_R_TRACK"]].tolist())) == list(range(8))
    venv.close()


def test_full_size_properties_config3():
    """BASELINE config 3 shape (8192 envs x 10 cars on talladega): size-independent invariants after a driving rollout."""
    import torch
    from nascargymnasium_b200.engine import Engine
    eng = Engine(8192, 10, tracks=["talladega"], auto_reset=True)
    eng.reset_host()
    last = torch.empty((81920, 38), dtype=torch.float32, device="cuda:0")
    eng.rollout(120, seed=9, mode=1, obs_last=last.view(-1))
    torch.cuda.synchronize()
    o = last.cpu().numpy()
    assert np.isfinite(o).all() and o.min() >= -1.0 and o.max() <= 1.0
    assert (o[:, 4] >= 0).all() and (o[:, 22:] >= 0).all()
    st = eng.read_stats()
    assert st["car_steps"] == 81920 * 120 and st["overflow"] == 0
    recs = eng.get_state_host()
    u = recs.view(np.uint32)
    steps = u[:, R["NCG_R_STEP"]].reshape(8192, 10)
    assert (steps == steps[:, :1]).all()                  # cars of an env share the env clock
    assert ((u[:, R["NCG_R_NCONTACT"]] & 255) <= L.MAX_CONTACTS).all()
    eng.close()


def test_vector_env_results_are_handed_out_without_copy_and_never_overwritten_while_held():
    """step() returns views of page-locked buffers the kernel wrote; a buffer is reused only after the caller drops them."""
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    E = 96
    venv = NascarVectorEnv(E, track_file="tracks/nascar.track")
    ref = NascarVectorEnv(E, track_file="tracks/nascar.track")
    venv.reset(); ref.reset()
    rng = np.random.default_rng(2)
    held = []
    for t in range(7):
        a = rng.uniform(-1, 1, (E, 2)).astype(np.float32)
        o, r, te, tr, _ = venv.step(a)
        o2, r2, te2, tr2, _ = ref.step(a)                # second env: its results are dropped every step (the ring stays at its initial three blocks)
        assert np.array_equal(o, o2) and np.array_equal(r, r2) and te.dtype == np.bool_ and o.dtype == np.float32
        held.append((o, o.copy(), r, r.copy(), te, te.copy()))
        del o2, r2, te2, tr2
    for o, oc, r, rc, te, tec in held:
        assert np.array_equal(o, oc) and np.array_equal(r, rc) and np.array_equal(te, tec)
    assert len(venv._ring) >= 7 and len(ref._ring) == 3
    # the mapped path and the staged-copy path are the same kernel: same numbers
    eng_obs, eng_rew, *_ = ref.engine.step_host(a)
    o3, r3, *_ = venv.step(a)
    assert np.array_equal(o3, eng_obs) and np.array_equal(r3, eng_rew)
    venv.close(); ref.close()


def test_rollout_is_deterministic_across_launch_shapes_on_the_track_mix():
    """The physics-warp / ray-warp hand-off is double-buffered across steps: one 600-step launch, launches of 7+250+343
    steps and five single steps + 595 must give bit-identical records and observations (config 4 track mix)."""
    import torch
    from nascargymnasium_b200.engine import Engine
    E, Tn = 2048, 600
    tid = (np.arange(E) * len(T.BUILTIN_TRACK_NAMES) // E).astype(np.int32)
    outs = []
    for chunks in ([600], [7, 250, 343], [1] * 5 + [595]):
        eng = Engine(E, 1, tracks=list(T.BUILTIN_TRACK_NAMES), auto_reset=True)
        eng.reset_host(track_id=tid)
        obs = torch.zeros((Tn, E, 38), dtype=torch.float32, device="cuda:0")
        rew = torch.zeros((Tn, E), dtype=torch.float32, device="cuda:0")
        t0 = 0
        for n in chunks:
            eng.rollout(n, seed=11, mode=1, obs_rollout=obs[t0:t0 + n].reshape(-1), reward_rollout=rew[t0:t0 + n].reshape(-1))
            t0 += n
        torch.cuda.synchronize()
        outs.append((eng.get_state_host(), obs.cpu().numpy(), rew.cpu().numpy(), eng.read_stats()))
        eng.close()
    ref = outs[0]
    assert ref[3]["contact_steps"] > 100 and ref[3]["episodes"] > 0           # the driving distribution hits walls and resets
    for o in outs[1:]:
        assert np.array_equal(o[0].view(np.uint32), ref[0].view(np.uint32))
        assert np.array_equal(o[1].view(np.uint32), ref[1].view(np.uint32)) and np.array_equal(o[2].view(np.uint32), ref[2].view(np.uint32))
        assert o[3]["car_steps"] == E * Tn == ref[3]["car_steps"]


def test_teacher_forced_steps_on_all_tracks_in_one_engine():
    """One engine holding all 8 tracks in blocks of envs (the CTA table keeps every CTA on one track): every block is
    teacher-forced from oracle states of its own track and must match the oracle like the single-track engines do."""
    from nascargymnasium_b200.engine import Engine
    per = 60
    names = list(T.BUILTIN_TRACK_NAMES)
    cases = [P.collect_cases(nm, per, kind="drive", seed=20 + i) for i, nm in enumerate(names)]
    n_each = min(len(c[0]) for c in cases)
    E = n_each * len(names)
    eng = Engine(E, 1, tracks=names, auto_reset=False)
    tid = np.repeat(np.arange(len(names), dtype=np.int32), n_each)
    eng.reset_host(track_id=tid)
    recs = np.concatenate([c[0][:n_each] for c in cases])
    recs.view(np.uint32)[:, R["NCG_R_TRACK"]] = tid
    eng.set_state_host(recs)
    raws = np.concatenate([np.array(c[2][:n_each], dtype=np.float32) for c in cases])
    obs, rew, te, tr, _ = eng.step_host(raws)
    got = eng.get_state_host()
    for i, nm in enumerate(names):
        sl = slice(i * n_each, (i + 1) * n_each)
        exp = {k: v[:n_each] for k, v in cases[i][3].items()}
        g = got[sl].copy()
        g.view(np.uint32)[:, R["NCG_R_TRACK"]] = exp["records"].view(np.uint32)[:, R["NCG_R_TRACK"]]     # single-track oracle: id 0
        bad, report = P.check_cases(g, obs[sl], rew[sl], te[sl], tr[sl], exp, label=nm)
        assert bad == 0, report
    eng.close()


def test_ppo_example_runs_on_device_end_to_end():
    """examples/ppo_rollout.py (the learn/ppo.py loop shape over NascarVectorEnv.step_torch): two small iterations."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "ppo_rollout.py"), "--envs", "256", "--n-steps", "32", "--iters", "2",
                          "--track", "martinsville", "--discrete", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 2 and lines[-1]["obs_device"].startswith("cuda")
    assert all(np.isfinite(l["policy_loss"]) and np.isfinite(l["value_loss"]) and l["rollout_env_steps_per_s"] > 0 for l in lines)


def test_engine_bounce_and_tunnelling_known_answers():
    """The analytic checks of tests/test_oracle_physics_kat.py on the CUDA engine: restitution 0.25 on a head-on hit,
    the summed impulse in obs[19], and TOI sub-stepping keeping a 110 m/s car out of a 1 m wall."""
    import math
    from nascargymnasium_b200.engine import Engine
    from oracle import oracle as O
    from tests import test_oracle_physics_kat as KAT
    speeds = [5.0, 20.0, 110.0]
    recs = []
    for v0 in speeds:
        env = O.OracleEnv(T.builtin_track_text("daytona"))
        env.reset()
        KAT._place(env, 50.0, 0.0, math.pi / 2, 0.0, v0)
        recs.append(P.oracle_to_record(env.get_state()))
    eng = Engine(len(speeds), 1, tracks=["daytona"], auto_reset=False)
    eng.reset_host()
    eng.set_state_host(np.array(recs, dtype=np.float32))
    zero = np.zeros((len(speeds), 2), dtype=np.float32)
    vy_prev = np.array(speeds)
    bounced = [None] * len(speeds)
    worst = np.full(len(speeds), -1e9)
    for _ in range(120):
        obs, rew, te, tr, _ = eng.step_host(zero)
        st = eng.get_state_host()
        y, vy = st[:, R["NCG_R_Y"]], st[:, R["NCG_R_VY"]]
        worst = np.maximum(worst, y + KAT.HALF_LEN - KAT.WALL_FACE_Y)
        for i in range(len(speeds)):
            if bounced[i] is None and vy[i] < 0:
                bounced[i] = (vy_prev[i], -vy[i], obs[i, 19] * 50000.0)
        vy_prev = vy.copy()
    eng.close()
    assert all(b is not None for b in bounced)
    assert (worst < 0.05).all(), worst                                   # never past the wall face, even at 110 m/s
    for v_in, v_out, imp in bounced[:2]:
        assert v_out == pytest.approx(0.25 * v_in, rel=0.03)
        assert imp == pytest.approx(min(1500.0 * 1.25 * v_in, 50000.0), rel=0.03)


def test_soak_driving_rollout_stays_finite_on_every_track():
    """8000 steps of the driving distribution on all 8 tracks (wall contacts, TOI events, disables, resets): state and
    observations stay finite and in range, and no per-car cap (contacts, manifolds, listener entries) overflows."""
    import torch
    from nascargymnasium_b200.engine import Engine
    E = 1024
    names = list(T.BUILTIN_TRACK_NAMES)
    eng = Engine(E, 1, tracks=names, auto_reset=True)
    eng.reset_host(track_id=(np.arange(E) * len(names) // E).astype(np.int32))
    last = torch.empty((E, 38), dtype=torch.float32, device="cuda:0")
    for i in range(8):
        eng.rollout(1000, seed=40 + i, mode=1, obs_last=last.view(-1))
    torch.cuda.synchronize()
    st = eng.read_stats()
    o, recs = last.cpu().numpy(), eng.get_state_host()
    assert st["car_steps"] == E * 8000 and st["contact_steps"] > 10000 and st["toi_events"] > 0 and st["episodes"] > 0
    assert st["overflow"] == 0
    assert np.isfinite(o).all() and o.min() >= -1.0 and o.max() <= 1.0
    assert np.isfinite(recs[:, [R["NCG_R_X"], R["NCG_R_Y"], R["NCG_R_ANGLE"], R["NCG_R_VX"], R["NCG_R_VY"], R["NCG_R_OMEGA"]]]).all()
    eng.close()


def test_sb3_vec_env_adapter_against_a_stub_of_the_sb3_base_class(monkeypatch):
    """stable_baselines3 is not in the image; a minimal stand-in for its VecEnv base class lets the adapter's own logic
    (step_async/step_wait, terminal_observation, TimeLimit.truncated, episode stats) run on the GPU engine."""
    import sys
    import types

    class VecEnv:                                            # the part of SB3's base class the adapter relies on
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

        def _get_indices(self, indices):
            return range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

    names = ["stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.vec_env",
             "stable_baselines3.common.vec_env.base_vec_env"]
    mods = {n: types.ModuleType(n) for n in names}
    mods[names[-1]].VecEnv = VecEnv
    for n, m in mods.items():
        monkeypatch.setitem(sys.modules, n, m)
    from nascargymnasium_b200.vector_env import make_sb3_vec_env
    env = make_sb3_vec_env(32, "tracks/martinsville.track", discrete_action_space=True)
    obs = env.reset()
    assert obs.shape == (32, 38) and env.num_envs == 32
    acts = np.zeros(32, dtype=np.int64)
    acts[::2] = 1
    for t in range(1, 601):
        obs, rew, done, infos = env.step(acts)
    assert done[1::2].all() and not done[::2].any()          # the idle cars are stuck-disabled at step 600
    assert infos[1]["terminal_observation"].shape == (38,) and infos[1]["episode"]["l"] == 600
    assert infos[1]["TimeLimit.truncated"] is False and infos[0] == {}
    assert env.env_is_wrapped(None) == [False] * 32 and len(env.get_attr("num_envs")) == 32
    # per-env answers for the attributes SB3's wrappers ask for
    assert env.get_attr("track_file", indices=[0, 5]) == ["tracks/martinsville.track"] * 2
    assert env.get_attr("render_mode") == [None] * 32 and env.get_attr("num_cars", 3) == [1]
    assert env.get_attr("action_space", [1])[0].n == 5
    info7 = env.env_method("get_info", indices=[7])[0]
    assert info7["num_cars"] == 1 and "lap_timing" in info7["cars"][0]
    assert env.seed(10)[:3] == [10, 11, 12]
    env.close()


def test_unstaged_track_table_gives_identical_results(monkeypatch):
    """The fallback for track tables too large to stage (table read through L1/L2 instead of shared memory) must be the
    same computation: bit-identical records and observations after a driving rollout."""
    import torch
    from nascargymnasium_b200.engine import Engine
    outs = []
    for no_stage in ("0", "1"):
        monkeypatch.setenv("NCG_NO_STAGE", no_stage)
        eng = Engine(300, 1, tracks=["trioval", "martinsville"], auto_reset=True)
        eng.reset_host(track_id=(np.arange(300) * 2 // 300).astype(np.int32))
        obs = torch.zeros((150, 300, 38), dtype=torch.float32, device="cuda:0")
        eng.rollout(150, seed=4, mode=1, obs_rollout=obs.view(-1))
        torch.cuda.synchronize()
        outs.append((eng.get_state_host(), obs.cpu().numpy()))
        eng.close()
    assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))
    assert np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))


def test_ray_queue_and_fixed_ray_mapping_give_identical_results(monkeypatch):
    """Large batches hand the rays of a CTA's cars out from a queue (longest first) and put two physics warps in a CTA,
    small ones use a fixed lane -> rays map and one physics warp; all run the same computation per car, so records and
    observations must be bit-identical -- staged and unstaged, with partly filled and unpaired groups, on two tracks."""
    import torch
    from nascargymnasium_b200.engine import Engine
    for E, C in ((333, 1), (700, 3)):
        _queue_variants_agree(monkeypatch, E, C)


def _queue_variants_agree(monkeypatch, E, C):
    import torch
    from nascargymnasium_b200.engine import Engine
    outs = []
    for queue, no_stage, rpl, pw, rw8 in (("0", "0", "2", "1", "0"), ("1", "0", "2", "1", "0"), ("1", "0", "4", "1", "0"), ("1", "1", "4", "1", "0"),
                                          ("0", "0", "4", "1", "0"), ("1", "0", "4", "2", "0"), ("1", "1", "4", "2", "0"), ("1", "0", "4", "2", "1"),
                                          ("1", "1", "4", "2", "1"), ("0", "0", "2", "4", "0"), ("1", "1", "2", "4", "0")):
        monkeypatch.setenv("NCG_PAIR_RW8", rw8)        # the pair shape with eight ray warps (96 registers) instead of six
        monkeypatch.setenv("NCG_RAY_QUEUE", queue)
        monkeypatch.setenv("NCG_NO_STAGE", no_stage)
        monkeypatch.setenv("NCG_RAYS_PER_LANE", rpl)
        monkeypatch.setenv("NCG_PHYS_WARPS", pw)       # 2: a CTA serves two groups of envs (two physics warps, six ray warps);
        # 4: a group's cars spread over four physics warps of eight lanes (single-car envs; ignored for C = 3)
        eng = Engine(E, C, tracks=["daytona", "nascar2"], auto_reset=True)
        eng.reset_host(track_id=(np.arange(E) * 2 // E).astype(np.int32))
        obs = torch.zeros((200, E * C, 38), dtype=torch.float32, device="cuda:0")
        eng.rollout(200, seed=9, mode=1, obs_rollout=obs.view(-1))
        torch.cuda.synchronize()
        outs.append((eng.get_state_host(), obs.cpu().numpy(), eng.read_stats()))
        eng.close()
    for o in outs[1:]:
        assert np.array_equal(o[0].view(np.uint32), outs[0][0].view(np.uint32))
        assert np.array_equal(o[1].view(np.uint32), outs[0][1].view(np.uint32))
        assert o[2]["ray_tests"] == outs[0][2]["ray_tests"]            # the same boxes were tested, only by other lanes


def test_full_size_config4_and_shard_independence():
    """BASELINE config 4 shape (65536 single-car envs over all 8 tracks in blocks): invariants after a driving rollout, and
    the property sharding relies on -- an env's trajectory does not depend on which other envs share the launch: the
    first 4096 envs of the big batch equal a 4096-env engine on the same track with the same Philox keys, bit for bit
    (different CTA shapes: 3 resident CTAs of 32 / 4 rays per lane vs one CTA of 28 per SM / 2 rays per lane)."""
    import torch
    from nascargymnasium_b200.engine import Engine
    names = list(T.BUILTIN_TRACK_NAMES)
    E, Tn = 65536, 300
    big = Engine(E, 1, tracks=names, auto_reset=True)
    big.reset_host(track_id=(np.arange(E) * len(names) // E).astype(np.int32))
    last = torch.empty((E, 38), dtype=torch.float32, device="cuda:0")
    big.rollout(Tn, seed=21, mode=1, obs_last=last.view(-1))
    torch.cuda.synchronize()
    st = big.read_stats()
    o = last.cpu().numpy()
    recs = big.get_state_host()
    big.close()
    assert st["car_steps"] == E * Tn and st["overflow"] == 0
    assert np.isfinite(o).all() and o.min() >= -1.0 and o.max() <= 1.0 and (o[:, 22:] >= 0).all()
    u = recs.view(np.uint32)
    assert (u[:, R["NCG_R_STEP"]] <= Tn).all()                 # (0 for an env that was auto-reset in the last step)
    assert sorted(set(u[:, R["NCG_R_TRACK"]].tolist())) == list(range(8))
    small = Engine(4096, 1, tracks=names, auto_reset=True)
    small.reset_host(track_id=np.zeros(4096, dtype=np.int32))
    last_s = torch.empty((4096, 38), dtype=torch.float32, device="cuda:0")
    small.rollout(Tn, seed=21, mode=1, obs_last=last_s.view(-1))
    torch.cuda.synchronize()
    recs_s = small.get_state_host()
    small.close()
    assert np.array_equal(recs_s.view(np.uint32), u[:4096])
    assert np.array_equal(last_s.cpu().numpy().view(np.uint32), o[:4096].view(np.uint32))
