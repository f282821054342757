"""Batched environments over the CUDA engine.

* :class:`NascarVectorEnv` -- Gymnasium-0.29 ``VectorEnv`` semantics (same-step auto-reset, ``final_observation``
  in ``info``) for E CarEnv instances of C cars; numpy in/out (``step``) or torch CUDA tensors in/out
  (``step_torch``, observations never leave the device).
* :func:`make_sb3_vec_env` -- Stable-Baselines3 ``VecEnv`` adapter (import-guarded: SB3 is not in the build image),
  the surface learn/ppo.py, sac.py and td3.py drive (/root/reference/learn/ppo.py:65-100).
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np

from . import constants as K
from . import info as I
from . import layout as L
from . import spaces as S
from . import track as T
from .engine import Engine


def _batched_space(space, n):
    if isinstance(space, S.Box):
        return S.Box(low=np.broadcast_to(space.low, (n,) + space.shape).copy(), high=np.broadcast_to(space.high, (n,) + space.shape).copy(),
                     shape=(n,) + space.shape, dtype=space.dtype)
    if isinstance(space, S.Discrete):
        return S.MultiDiscrete([space.n] * n)
    return S.MultiDiscrete(np.broadcast_to(space.nvec, (n,) + space.nvec.shape).copy())


class StepInfo(dict):
    """`info` of a step in which at least one env finished.  The terminal observations arrive from the engine as a compact
    (n_done, [C,] 38) array plus the indices of the finished envs; the Gymnasium-0.29 object array ``final_observation``
    (and the Gymnasium-1.x dense ``final_obs``) are built from them only when somebody asks, because with thousands of
    envs and short episodes building one Python object per finished env every step costs more than the step itself."""

    def __init__(self, done: np.ndarray, index: np.ndarray, rows: np.ndarray, obs_shape, episode: dict):
        super().__init__(_final_observation=done, _final_obs=done, _episode=done, episode=episode,
                         final_obs_index=index, final_obs_rows=rows)
        self._obs_shape = obs_shape

    def __missing__(self, key):
        if key == "final_observation":
            fo = np.empty(self._obs_shape[0], dtype=object)
            rows = self["final_obs_rows"]
            for j, e in enumerate(self["final_obs_index"]):
                fo[e] = rows[j]
            self[key] = fo
            return fo
        if key == "final_obs":
            dense = np.zeros(self._obs_shape, dtype=np.float32)
            dense[self["final_obs_index"]] = self["final_obs_rows"]
            self[key] = dense
            return dense
        raise KeyError(key)

    def __contains__(self, key):
        return key in ("final_observation", "final_obs") or dict.__contains__(self, key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class NascarVectorEnv:
    metadata = {"render_modes": ["human"], "render_fps": 60, "autoreset_mode": "same_step"}

    def __init__(self, num_envs: int, track_file: Union[None, str, Sequence[str]] = None, num_cars: int = 1,
                 discrete_action_space: bool = False, reset_on_lap: bool = False, device: int = 0, track_info: bool = False,
                 max_result_blocks: int = 64, validate_actions: bool = True, redraw_tracks: bool = True,
                 car_contacts: bool = False, start_grid=(8.0, 3.0)):
        self.validate_actions = bool(validate_actions)    # assert action_space.contains, like the reference's step()
        self.max_result_blocks = int(max_result_blocks)   # result buffers kept alive for the caller before step() copies
        if num_cars < 1 or num_cars > K.MAX_CARS:
            raise ValueError(f"Number of cars must be between 1 and {K.MAX_CARS}")
        # track_file=None is the reference's random-track mode (car_env.py:243-303; learn/ppo.py:65-77 trains that way): an env
        # draws a track at reset() and, when an episode ends, restarts on ANOTHER one.  redraw_tracks=False keeps every env on
        # the track it drew at reset() (needed to capture step_torch in a CUDA graph: re-grouping envs is host work).
        self.random_tracks = track_file is None
        if track_file is None:
            tracks = [f"tracks/{n}.track" for n in T.BUILTIN_TRACK_NAMES]
        elif isinstance(track_file, str):
            tracks = [track_file]
        else:
            tracks = list(track_file)
        for t in tracks:
            T.load_track(t)
        self.num_envs, self.num_cars, self.discrete = int(num_envs), int(num_cars), bool(discrete_action_space)
        self.tracks = tracks
        self.single_action_space, self.single_observation_space = S.make_spaces(discrete_action_space, num_cars)
        self.action_space = _batched_space(self.single_action_space, num_envs)
        self.observation_space = _batched_space(self.single_observation_space, num_envs)
        self.engine = Engine(num_envs, num_cars, tracks=tracks, discrete=discrete_action_space, reset_on_lap=reset_on_lap,
                             auto_reset=True, device=device, track_info=track_info, car_contacts=car_contacts, grid=start_grid)
        self.redraw_tracks = bool(redraw_tracks) and self.random_tracks and len(tracks) > 1
        # a list of tracks: equal blocks of envs per track; random mode: drawn per env at reset()
        self._initial_track_id = (np.arange(num_envs, dtype=np.int64) * len(tracks) // num_envs).astype(np.int32)
        self._obs_shape = (num_envs, K.OBS_DIM) if num_cars == 1 else (num_envs, num_cars, K.OBS_DIM)
        self._rew_shape = (num_envs,) if num_cars == 1 else (num_envs, num_cars)
        self._aux, self._ring, self._ring_pos, self._spill = None, [], 0, None
        self._torch_bufs = None
        self._fast = None
        self._posted = None
        self.closed = False

    # ------------------------------------------------------------------ numpy API
    @property
    def track_id(self) -> np.ndarray:
        """Current env -> track index (into self.tracks); in random-track mode it changes as episodes end."""
        return self.engine.env_tracks()

    def _draw_tracks(self, seed):
        """reset(): every env of a random-track batch draws its own track (CarEnv._select_random_track); `seed` makes the
        draw and the later re-draws reproducible (the reference seeds from pid + clock, i.e. not at all)."""
        if not self.random_tracks:
            return self._initial_track_id
        rng = np.random.default_rng(seed)
        self.engine.set_track_redraw(self.redraw_tracks, int(rng.integers(0, 2 ** 63)))
        return rng.integers(0, len(self.tracks), size=self.num_envs).astype(np.int32)

    def reset(self, seed=None, options=None):
        obs = self.engine.reset_host(track_id=self._draw_tracks(seed), fresh=True)
        return obs.reshape(self._obs_shape), {}

    def _free_result_block(self, exclude=None):
        """The next block of the ring nobody holds views of any more (and that is not `exclude`), or None."""
        for _ in range(len(self._ring)):
            self._ring_pos = (self._ring_pos + 1) % len(self._ring)
            blk = self._ring[self._ring_pos]
            if blk is not exclude and not blk.busy():
                return blk
        return None

    def _next_result_block(self):
        """Result buffers nobody holds views of any more: the next one of the ring, or a new one while the caller keeps
        earlier results alive (so returned arrays are never overwritten -- copy semantics without the copy)."""
        blk = self._free_result_block()
        if blk is not None:
            return blk
        if len(self._ring) >= self.max_result_blocks:
            return None
        blk = self._new_result_block()
        self._ring.insert(self._ring_pos + 1, blk)
        self._ring_pos += 1
        return blk

    def _new_result_block(self):
        """Result buffers with the views step() hands out and the argument block of the C call, both made once."""
        import ctypes
        from .engine import MappedBuffers
        blk = self.engine.result_block()
        r = blk.arrays
        blk.cache_views((r["obs"].reshape(self._obs_shape), r["reward"].reshape(self._rew_shape),
                         r["terminated"].view(np.bool_), r["truncated"].view(np.bool_)))
        a, p = self._aux.ptrs, blk.ptrs
        blk.cbuf = MappedBuffers(a["actions"].value, p["obs"].value, p["reward"].value, p["terminated"].value, p["truncated"].value,
                                 a["final_obs"].value, a["ep_return"].value, a["ep_length"].value)
        blk.cref = ctypes.byref(blk.cbuf)
        return blk

    def step(self, actions):
        """actions: (E[,C],2) float32 in [-1,1] or (E[,C]) ints.  Host buffers in, host buffers out.  The actions are
        checked and copied into a page-locked buffer the kernel reads directly (one pass, in the library); observations,
        rewards and flags are written by the kernel straight into page-locked result buffers, which are handed out without a
        copy and not reused while the caller still references them.  step() is step_async() + step_wait() in one body (the two
        calls and the hand-over between them cost 0.7 us of a 35 us step)."""
        if self._aux is None or self._posted is not None:
            self.step_async(actions)                # (first call: buffers are made; or a misuse that step_async reports)
            return self.step_wait()
        aux = self._aux.arrays
        a = actions if type(actions) is np.ndarray else np.asarray(actions)
        if a.dtype != self._act_dtype or not a.flags.c_contiguous or a.size != aux["actions"].size:
            if self.validate_actions:
                if self.discrete:
                    assert a.dtype.kind in "iu" and a.size == aux["actions"].size and a.min() >= 0 and a.max() < 5, "Invalid action"
                else:
                    assert a.dtype == np.float32 and a.size == aux["actions"].size, "Invalid action"
            a = np.ascontiguousarray(a, dtype=self._act_dtype).reshape(aux["actions"].shape)
        blk, self._next_blk = self._next_blk, None
        if blk is None:
            blk = self._next_result_block()
        spill = blk is None
        if spill:
            if self._spill is None:
                self._spill = self._new_result_block()
            blk = self._spill
        h = self.engine._h
        rc = self._post_c(h, a.__array_interface__["data"][0], 1 if self.validate_actions else 0, blk.cref)
        if rc:
            if self.engine._lib.ncg_last_error() == b"Invalid action":
                raise AssertionError("Invalid action")
            from .engine import _check
            _check(rc)
        self._next_blk = self._free_result_block(exclude=blk)       # (while the GPU steps)
        rc = self._wait_c(h, self._done_ref)
        if rc:
            from .engine import _check
            _check(rc)
        obs, rew, te, tr = blk.views
        if spill:
            obs, rew, te, tr = obs.copy(), rew.copy(), te.copy(), tr.copy()
        if not self._done_flag.value:
            return obs, rew, te, tr, {}
        return obs, rew, te, tr, self._episode_info(te, tr)

    def _episode_info(self, te, tr):
        aux = self._aux.arrays
        done = te | tr
        idx = np.flatnonzero(done)
        rows = aux["final_obs"].reshape(self._obs_shape)[idx]          # one gather-copy out of the mapped buffer
        ep_r = np.zeros(self._rew_shape, dtype=np.float64)
        ep_l = np.zeros(self.num_envs, dtype=np.int64)
        ep_r[idx] = aux["ep_return"].reshape(self._rew_shape)[idx]
        ep_l[idx] = aux["ep_length"][idx]
        return StepInfo(done.copy(), idx, rows, self._obs_shape, {"r": ep_r, "l": ep_l})

    def step_async(self, actions):
        """Hand a step to the GPU and return at once (ncg_step_mapped_post; the interface of stable-baselines3's VecEnv, which
        /root/reference/learn/ppo.py:65-78 drives).  Whatever the caller does before step_wait() overlaps the step.  `actions` is
        copied before this returns.  Calling anything else of the env in between completes the step first."""
        if self._aux is None:
            import ctypes
            self._aux = self.engine.aux_block()
            self._ring, self._ring_pos = [self._new_result_block() for _ in range(3)], 0
            self._act_dtype = np.dtype(np.int32 if self.discrete else np.float32)
            self._done_flag = ctypes.c_int32(0)
            self._done_ref = ctypes.byref(self._done_flag)
            self._post_c, self._wait_c = self.engine._lib.ncg_step_mapped_post, self.engine._lib.ncg_step_mapped_wait
            self._next_blk, self._posted = None, None
        if self._posted is not None:
            raise RuntimeError("step_async() called twice without step_wait()")
        aux = self._aux.arrays
        a = actions if type(actions) is np.ndarray else np.asarray(actions)
        if a.dtype != self._act_dtype or not a.flags.c_contiguous or a.size != aux["actions"].size:
            # CarEnv.step asserts action_space.contains(action) on every step (/root/reference/src/car_env.py:694)
            if self.validate_actions:
                if self.discrete:
                    assert a.dtype.kind in "iu" and a.size == aux["actions"].size and a.min() >= 0 and a.max() < 5, "Invalid action"
                else:
                    assert a.dtype == np.float32 and a.size == aux["actions"].size, "Invalid action"
            a = np.ascontiguousarray(a, dtype=self._act_dtype).reshape(aux["actions"].shape)
        blk, self._next_blk = self._next_blk, None
        if blk is None:
            blk = self._next_result_block()
        spill = blk is None
        if spill:                                   # the caller holds max_result_blocks results: fall back to copying
            if self._spill is None:
                self._spill = self._new_result_block()
            blk = self._spill
        rc = self._post_c(self.engine._h, a.__array_interface__["data"][0], 1 if self.validate_actions else 0, blk.cref)
        if rc:
            if self.engine._lib.ncg_last_error() == b"Invalid action":
                raise AssertionError("Invalid action")
            from .engine import _check
            _check(rc)
        self._posted = (blk, spill)
        # (while the GPU steps) the block of the next step, among those that exist: a block that is free now stays free, the caller
        # can only get hold of the one being filled
        self._next_blk = self._free_result_block(exclude=blk)

    def step_wait(self):
        """The results of the step handed over by step_async()."""
        if getattr(self, "_posted", None) is None:
            raise RuntimeError("step_wait() without step_async()")
        (blk, spill), self._posted = self._posted, None
        rc = self._wait_c(self.engine._h, self._done_ref)
        if rc:
            from .engine import _check
            _check(rc)
        obs, rew, te, tr = blk.views
        if spill:
            obs, rew, te, tr = obs.copy(), rew.copy(), te.copy(), tr.copy()
        return obs, rew, te, tr, (self._episode_info(te, tr) if self._done_flag.value else {})

    # ------------------------------------------------------------------ torch API (device-resident)
    def _bufs(self):
        if self._torch_bufs is None:
            import torch
            dev = f"cuda:{self.engine.device}"
            N, E = self.engine.num_cars, self.num_envs
            self._torch_bufs = dict(obs=torch.empty((N, K.OBS_DIM), dtype=torch.float32, device=dev),
                                    final=torch.zeros((N, K.OBS_DIM), dtype=torch.float32, device=dev),
                                    rew=torch.empty(N, dtype=torch.float32, device=dev),
                                    te=torch.empty(E, dtype=torch.uint8, device=dev), tr=torch.empty(E, dtype=torch.uint8, device=dev),
                                    ep_return=torch.zeros(N, dtype=torch.float32, device=dev), ep_length=torch.zeros(E, dtype=torch.int32, device=dev),
                                    any_done=torch.zeros(1, dtype=torch.int32, device=dev))
            b = self._torch_bufs
            # Monitor-style episode statistics stay on the device too (learn/ppo.py:69 wraps every env in a Monitor)
            self.engine.set_episode_outputs(b["ep_return"], b["ep_length"], b["any_done"])
        return self._torch_bufs

    @property
    def episode_returns(self):
        """CUDA tensor (E[,C]): return of the episode an env finished most recently (valid where terminated | truncated
        was set by a step_torch call; rows of running envs keep their previous value)."""
        return self._bufs()["ep_return"].view(self._rew_shape)

    @property
    def episode_lengths(self):
        """CUDA int32 tensor (E,): length in steps of the episode an env finished most recently."""
        return self._bufs()["ep_length"]

    def reset_torch(self, seed=None):
        import torch
        b = self._bufs()
        tid = torch.as_tensor(self._draw_tracks(seed), device=b["obs"].device)
        self.engine.reset(obs=b["obs"].view(-1), track_id=tid, fresh=True)
        return b["obs"].view(self._obs_shape)

    def step_torch(self, actions):
        """actions: CUDA tensor float32 (E[,C],2) or int32 (E[,C]).  Returns views of internal CUDA buffers
        (obs, reward, terminated, truncated, final_obs) that are overwritten by the next call.  Nothing synchronises (except
        in random-track mode, see ncg_set_track_redraw); the call costs a few microseconds of host time: the output
        pointers and views are prepared once, only the action tensor is checked per call."""
        f = self._fast
        if f is None:
            import ctypes
            import torch
            b = self._bufs()
            eng = self.engine
            f = self._fast = dict(
                ptrs=[ctypes.c_void_p(b[k].data_ptr()) for k in ("obs", "rew", "te", "tr", "final")],
                out=(b["obs"].view(self._obs_shape), b["rew"].view(self._rew_shape), b["te"], b["tr"], b["final"].view(self._obs_shape)),
                dtype=torch.int32 if self.discrete else torch.float32, numel=eng.num_cars * (1 if self.discrete else 2),
                stream=torch.cuda.current_stream, vp=ctypes.c_void_p, step=eng._lib.ncg_step, h=eng._h, dev=eng.device)
        if not (actions.is_cuda and actions.dtype == f["dtype"] and actions.numel() == f["numel"] and actions.is_contiguous()
                and actions.device.index == f["dev"]):
            raise ValueError(f"actions: expected a contiguous {f['dtype']} CUDA tensor with {f['numel']} elements on cuda:{f['dev']}")
        vp, p = f["vp"], f["ptrs"]
        rc = f["step"](f["h"], vp(actions.data_ptr()), p[0], p[1], p[2], p[3], p[4], vp(f["stream"](f["dev"]).cuda_stream))
        if rc:
            from .engine import _check
            _check(rc)
        return f["out"]

    # ------------------------------------------------------------------ info on demand
    def get_info(self, env_index: int) -> dict:
        recs = self.engine.get_state_host().reshape(self.num_envs, self.num_cars, L.RECORD_WORDS)
        return I.env_info(recs[env_index])

    def close(self):
        if not self.closed:
            # blocks the caller still references stay allocated until those arrays die (HostBlock.__del__ frees them)
            self._ring, self._aux, self._spill, self._next_blk, self._posted = [], None, None, None, None
            self.engine.close()
            self.closed = True


def make_sb3_vec_env(num_envs: int, track_file=None, discrete_action_space: bool = False, reset_on_lap: bool = False, device: int = 0):
    """Stable-Baselines3 ``VecEnv`` over :class:`NascarVectorEnv` (single-car envs).  Requires stable_baselines3."""
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv      # import-guarded: not in the offline image

    venv = NascarVectorEnv(num_envs, track_file=track_file, num_cars=1, discrete_action_space=discrete_action_space,
                           reset_on_lap=reset_on_lap, device=device)

    class _SB3(VecEnv):
        def __init__(self):
            super().__init__(num_envs, venv.single_observation_space, venv.single_action_space)
            self._actions = None

        def reset(self):
            return venv.reset(seed=getattr(self, "_seed", None))[0]

        def step_async(self, actions):
            venv.step_async(actions)          # the GPU starts on the step now; SB3 calls step_wait() next

        def step_wait(self):
            obs, rew, te, tr, info = venv.step_wait()
            done = te | tr
            infos = [{} for _ in range(num_envs)]
            if info:
                for j, e in enumerate(info["final_obs_index"]):
                    infos[e] = {"terminal_observation": info["final_obs_rows"][j], "TimeLimit.truncated": bool(tr[e] and not te[e]),
                                "episode": {"r": float(info["episode"]["r"][e]), "l": int(info["episode"]["l"][e])}}
            return obs, rew, done, infos

        def close(self):
            venv.close()

        # Per-env attributes the wrappers ask a (Subproc)VecEnv for.  There is no per-env Python object here, so the ones
        # SB3's Monitor / VecNormalize / evaluation helpers read are answered per env; anything else is the batched env's
        # own attribute, once per requested index (documented deviation: it is the same object for every index).
        _PER_ENV = {"render_mode": None, "num_cars": 1, "discrete_action_space": discrete_action_space, "reset_on_lap": reset_on_lap}

        def get_attr(self, attr_name, indices=None):
            idx = list(self._get_indices(indices))
            if attr_name == "track_file":
                tid = venv.track_id
                return [venv.tracks[int(tid[i])] for i in idx]
            if attr_name in ("action_space", "observation_space"):
                sp = venv.single_action_space if attr_name == "action_space" else venv.single_observation_space
                return [sp for _ in idx]
            if attr_name in self._PER_ENV:
                return [self._PER_ENV[attr_name] for _ in idx]
            if attr_name == "spec":
                return [None for _ in idx]
            return [getattr(venv, attr_name) for _ in idx]

        def set_attr(self, attr_name, value, indices=None):
            setattr(venv, attr_name, value)

        def env_method(self, method_name, *args, indices=None, **kwargs):
            idx = list(self._get_indices(indices))
            if method_name == "get_info":
                return [venv.get_info(i) for i in idx]
            out = getattr(venv, method_name)(*args, **kwargs)
            return [out for _ in idx]

        def env_is_wrapped(self, wrapper_class, indices=None):
            return [False] * len(self._get_indices(indices))

        def seed(self, seed=None):
            self._seed = seed
            return [None if seed is None else seed + i for i in range(num_envs)]

    return _SB3()
