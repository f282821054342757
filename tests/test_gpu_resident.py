"""GPU tests of the resident step kernel behind ncg_step_mapped (csrc/ncg_b200_res.cu): a host-driven loop stepped through the
mailbox gives bit for bit what per-step launches give, through idle exits, result-slot rotation, interleaved state reads and
every other entry point that has to end the resident launch first."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _venv(resident, **kw):
    from nascargymnasium_b200.vector_env import NascarVectorEnv
    old = os.environ.get("NCG_RESIDENT")
    os.environ["NCG_RESIDENT"] = "1" if resident else "0"
    try:
        return NascarVectorEnv(**kw)
    finally:
        if old is None:
            os.environ.pop("NCG_RESIDENT", None)
        else:
            os.environ["NCG_RESIDENT"] = old


def _actions(rng, v, drive=False):
    shape = (v.num_envs,) if v.num_cars == 1 else (v.num_envs, v.num_cars)
    if v.discrete:
        return rng.integers(0, 5, size=shape).astype(np.int64)
    a = rng.uniform(-1, 1, size=shape + (2,)).astype(np.float32)
    if drive:
        a[..., 0] = np.abs(a[..., 0]) * 0.8 + 0.2
        a[..., 1] = a[..., 1] * 0.4 + 0.2
    return a


def _same(r1, r2):
    for x, y in zip(r1[:4], r2[:4]):
        assert np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8))
    i1, i2 = r1[4], r2[4]
    assert bool(i1) == bool(i2)
    if i1:
        assert np.array_equal(i1["final_obs_index"], i2["final_obs_index"])
        assert np.array_equal(i1["final_obs_rows"].view(np.uint32), i2["final_obs_rows"].view(np.uint32))
        assert np.array_equal(i1["episode"]["r"], i2["episode"]["r"]) and np.array_equal(i1["episode"]["l"], i2["episode"]["l"])


@pytest.mark.parametrize("kw,steps,drive", [
    (dict(num_envs=4096, track_file="tracks/daytona.track"), 400, True),                      # one CTA per SM: <2,1,1>
    (dict(num_envs=8192, track_file="tracks/martinsville.track"), 150, True),                 # two per SM: <4,2,1>
    (dict(num_envs=300, track_file="tracks/talladega.track", num_cars=3), 150, False),        # multi-car envs, auto-reset within steps
    (dict(num_envs=512, track_file="tracks/nascar2.track", discrete_action_space=True), 150, False),
])
def test_resident_steps_equal_launched_steps(kw, steps, drive):
    va, vb = _venv(False, **kw), _venv(True, **kw)
    va.reset(seed=3); vb.reset(seed=3)
    rng = np.random.default_rng(5)
    l0 = vb.engine.launch_count
    # (one env after the other: a kernel launched for `va` would have to wait for the resident one to give its registers back)
    acts = [_actions(rng, va, drive) for t in range(steps)]
    keep = lambda r: tuple(np.array(x) for x in r[:4]) + (r[4],)        # (copies: the envs keep rotating two result blocks)
    ra = [keep(va.step(a)) for a in acts]
    rb = [keep(vb.step(a)) for a in acts]
    assert vb.engine.launch_count - l0 <= 3, "the resident kernel is launched once, not per step"
    for x, y in zip(ra, rb):
        _same(x, y)
    del ra, rb, x, y
    assert np.array_equal(va.engine.get_state_host().view(np.uint32), vb.engine.get_state_host().view(np.uint32))
    # ... and on from the state the resident launch left in HBM
    for t in range(20):
        a = _actions(rng, va, drive)
        _same(va.step(a), vb.step(a))
    sa, sb = va.engine.read_stats(), vb.engine.read_stats()
    assert sa == sb
    va.close(); vb.close()


def test_resident_kernel_leaves_when_idle_and_comes_back():
    os.environ["NCG_RESIDENT_IDLE_US"] = "2000"
    try:
        va, vb = _venv(False, num_envs=1024, track_file="tracks/martinsville.track"), _venv(True, num_envs=1024, track_file="tracks/martinsville.track")
    finally:
        os.environ.pop("NCG_RESIDENT_IDLE_US")
    va.reset(); vb.reset()
    rng = np.random.default_rng(1)
    l0 = vb.engine.launch_count
    for t in range(30):
        a = _actions(rng, va)
        _same(va.step(a), vb.step(a))
        if t % 10 == 9:
            time.sleep(0.02)                       # ten idle times: the kernel has left
    assert 3 <= vb.engine.launch_count - l0 <= 5
    va.close(); vb.close()


def test_resident_result_slots_rotate_and_wrap():
    """a caller that keeps results alive makes the env hand out fresh result blocks: more of them than the mailbox's slot table holds"""
    va = _venv(False, num_envs=256, track_file="tracks/daytona.track")
    vb = _venv(True, num_envs=256, track_file="tracks/daytona.track")
    vb.max_result_blocks = 40
    va.max_result_blocks = 40
    va.reset(); vb.reset()
    rng = np.random.default_rng(2)
    keep_a, keep_b = [], []
    for t in range(60):
        a = _actions(rng, va)
        ra, rb = va.step(a), vb.step(a)
        _same(ra, rb)
        keep_a.append(ra); keep_b.append(rb)
        if t == 45:
            keep_a.clear(); keep_b.clear()
    for ra, rb in zip(keep_a, keep_b):                 # held results were never overwritten
        _same(ra, rb)
    va.close(); vb.close()


def test_other_entry_points_end_the_resident_launch():
    import torch
    v = _venv(True, num_envs=2048, track_file="tracks/daytona.track")
    w = _venv(False, num_envs=2048, track_file="tracks/daytona.track")
    v.reset(); w.reset()
    rng = np.random.default_rng(4)
    for t in range(40):
        a = _actions(rng, v, True)
        _same(w.step(a), v.step(a))
        if t % 8 == 3:
            s = v.engine.get_state_host(); v.engine.set_state_host(s)
            assert np.array_equal(s.view(np.uint32), w.engine.get_state_host().view(np.uint32))
        if t % 8 == 5:
            ta = torch.as_tensor(a, device="cuda")
            ov, ow = v.step_torch(ta), w.step_torch(ta)
            torch.cuda.synchronize()
            assert torch.equal(ov[0], ow[0]) and torch.equal(ov[1], ow[1])
        if t == 30:
            v.reset(); w.reset()
    v.close(); w.close()


def test_a_slow_caller_falls_back_to_per_step_launches():
    """three idle exits in a row: the library stops keeping a kernel resident for a caller that steps slower than the idle time"""
    os.environ["NCG_RESIDENT_IDLE_US"] = "500"
    try:
        va, vb = _venv(False, num_envs=512, track_file="tracks/daytona.track"), _venv(True, num_envs=512, track_file="tracks/daytona.track")
    finally:
        os.environ.pop("NCG_RESIDENT_IDLE_US")
    va.reset(); vb.reset()
    rng = np.random.default_rng(7)
    l0 = vb.engine.launch_count
    for t in range(14):
        a = _actions(rng, va)
        _same(va.step(a), vb.step(a))
        time.sleep(0.004)
    # 3 resident launches that idled out, then one launch per step
    assert vb.engine.launch_count - l0 >= 3 + 10
    assert vb.engine.resident_stats["steps"] == 3
    va.close(); vb.close()


def test_post_and_wait_contract_through_the_c_abi():
    """one wait per post; a second post, or a wait without a post, is refused; any other entry point completes a posted step first"""
    import ctypes
    from nascargymnasium_b200.engine import Engine, MappedBuffers, NcgError
    E = 512
    for resident in (True, False):
        os.environ["NCG_RESIDENT"] = "1" if resident else "0"
        try:
            eng = Engine(E, 1, tracks=["daytona"], auto_reset=True)
            ref = Engine(E, 1, tracks=["daytona"], auto_reset=True)
        finally:
            os.environ.pop("NCG_RESIDENT")
        eng.reset_host(); ref.reset_host()
        aux, res = eng.aux_block(), eng.result_block()
        a, r = aux.ptrs, res.ptrs
        buf = MappedBuffers(a["actions"].value, r["obs"].value, r["reward"].value, r["terminated"].value, r["truncated"].value,
                            a["final_obs"].value, a["ep_return"].value, a["ep_length"].value)
        lib, h = eng._lib, eng._h
        done = ctypes.c_int32(0)
        assert lib.ncg_step_mapped_wait(h, ctypes.byref(done)) != 0                      # nothing posted
        rng = np.random.default_rng(3)
        acts = rng.uniform(-1, 1, (3, E, 2)).astype(np.float32)
        assert lib.ncg_step_mapped_post(h, acts[0].ctypes.data, 1, ctypes.byref(buf)) == 0
        assert lib.ncg_step_mapped_post(h, acts[1].ctypes.data, 1, ctypes.byref(buf)) != 0   # one wait per post
        assert lib.ncg_step_mapped_wait(h, ctypes.byref(done)) == 0
        o_ref = ref.step_host(acts[0])[0]
        assert np.array_equal(res.arrays["obs"], o_ref)
        bad = acts[1].copy(); bad[7, 1] = np.nan
        assert lib.ncg_step_mapped_post(h, bad.ctypes.data, 1, ctypes.byref(buf)) != 0 and lib.ncg_last_error() == b"Invalid action"
        assert np.array_equal(res.arrays["obs"], o_ref)                                  # nothing was stepped
        # a posted step that is never waited for: the state read completes it first (resident mode)
        assert lib.ncg_step_mapped_post(h, acts[1].ctypes.data, 1, ctypes.byref(buf)) == 0
        if resident:
            s1 = eng.get_state_host()
            ref.step_host(acts[1])
            assert np.array_equal(s1.view(np.uint32), ref.get_state_host().view(np.uint32))
        else:
            assert lib.ncg_step_mapped_wait(h, ctypes.byref(done)) == 0
        eng.close(); ref.close()


def test_step_async_and_step_wait_equal_step():
    """SB3's VecEnv interface on the batched env: host work between the two halves overlaps the step, results are those of step()"""
    va = _venv(True, num_envs=1024, track_file="tracks/daytona.track")
    vb = _venv(True, num_envs=1024, track_file="tracks/daytona.track")
    va.reset(); vb.reset()
    rng = np.random.default_rng(9)
    acts = [_actions(rng, va, True) for _ in range(300)]
    keep = lambda r: tuple(np.array(x) for x in r[:4]) + (r[4],)
    ra = [keep(va.step(a)) for a in acts]
    rb = []
    for a in acts:
        vb.step_async(a)
        a[...] = 0                                   # the actions were copied before step_async returned
        with pytest.raises(RuntimeError):
            vb.step_async(a)
        rb.append(keep(vb.step_wait()))
    with pytest.raises(RuntimeError):
        vb.step_wait()
    for x, y in zip(ra, rb):
        _same(x, y)
    va.close(); vb.close()
