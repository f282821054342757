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


class NascarVectorEnv:
    metadata = {"render_modes": ["human"], "render_fps": 60, "autoreset_mode": "same_step"}

    def __init__(self, num_envs: int, track_file: Union[None, str, Sequence[str]] = None, num_cars: int = 1,
                 discrete_action_space: bool = False, reset_on_lap: bool = False, device: int = 0, track_info: bool = False,
                 copy: bool = True):
        self.copy = copy            # False: step() returns views of the pinned staging buffers (valid until the next step)
        if num_cars < 1 or num_cars > K.MAX_CARS:
            raise ValueError(f"Number of cars must be between 1 and {K.MAX_CARS}")
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
                             auto_reset=True, device=device, track_info=track_info)
        # envs sorted by track id so every CTA serves one track
        self.track_id = (np.arange(num_envs, dtype=np.int64) * len(tracks) // num_envs).astype(np.int32)
        self._obs_shape = (num_envs, K.OBS_DIM) if num_cars == 1 else (num_envs, num_cars, K.OBS_DIM)
        self._rew_shape = (num_envs,) if num_cars == 1 else (num_envs, num_cars)
        self._ep_len = np.zeros(num_envs, dtype=np.int64)
        self._ep_ret = np.zeros(self._rew_shape, dtype=np.float64)
        self._torch_bufs = None
        self.closed = False

    # ------------------------------------------------------------------ numpy API
    def reset(self, seed=None, options=None):
        obs = self.engine.reset_host(track_id=self.track_id, fresh=True)
        self._ep_len[:] = 0
        self._ep_ret[...] = 0.0
        return obs.reshape(self._obs_shape), {}

    def step(self, actions):
        """actions: (E[,C],2) float32 in [-1,1] or (E[,C]) ints.  Host buffers in, host buffers out: the actions go through
        the library's page-locked staging buffer, one packed D2H copy brings obs/reward/flags back."""
        v = self.engine.pinned_views()
        v["actions"][...] = np.asarray(actions).reshape(v["actions"].shape)
        any_done = self.engine.step_pinned(want_final=True)
        obs = v["obs"].reshape(self._obs_shape)
        rew = v["reward"].reshape(self._rew_shape)
        if self.copy:
            obs, rew = obs.copy(), rew.copy()
        te, tr = v["terminated"].astype(bool), v["truncated"].astype(bool)
        self._ep_len += 1
        self._ep_ret += rew
        info = {}
        if any_done:
            done = te | tr
            fin = v["final_obs"].reshape(self._obs_shape)
            fo = np.empty(self.num_envs, dtype=object)
            ep_r = np.zeros(self._rew_shape, dtype=np.float64)
            ep_l = np.zeros(self.num_envs, dtype=np.int64)
            for e in np.nonzero(done)[0]:
                fo[e] = fin[e].copy()
            ep_r[done], ep_l[done] = self._ep_ret[done], self._ep_len[done]
            info = {"final_observation": fo, "_final_observation": done.copy(),
                    "episode": {"r": ep_r, "l": ep_l}, "_episode": done.copy()}
            self._ep_ret[done] = 0.0
            self._ep_len[done] = 0
        return obs, rew, te, tr, info

    # ------------------------------------------------------------------ torch API (device-resident)
    def _bufs(self):
        if self._torch_bufs is None:
            import torch
            dev = f"cuda:{self.engine.device}"
            N, E = self.engine.num_cars, self.num_envs
            self._torch_bufs = dict(obs=torch.empty((N, K.OBS_DIM), dtype=torch.float32, device=dev),
                                    final=torch.empty((N, K.OBS_DIM), dtype=torch.float32, device=dev),
                                    rew=torch.empty(N, dtype=torch.float32, device=dev),
                                    te=torch.empty(E, dtype=torch.uint8, device=dev), tr=torch.empty(E, dtype=torch.uint8, device=dev))
        return self._torch_bufs

    def reset_torch(self):
        import torch
        b = self._bufs()
        tid = torch.as_tensor(self.track_id, device=b["obs"].device)
        self.engine.reset(obs=b["obs"].view(-1), track_id=tid, fresh=True)
        return b["obs"].view(self._obs_shape)

    def step_torch(self, actions):
        """actions: CUDA tensor float32 (E[,C],2) or int32 (E[,C]).  Returns views of internal CUDA buffers
        (obs, reward, terminated, truncated, final_obs) that are overwritten by the next call."""
        b = self._bufs()
        self.engine.step(actions.contiguous().view(-1), b["obs"].view(-1), b["rew"], b["te"], b["tr"], b["final"].view(-1))
        return b["obs"].view(self._obs_shape), b["rew"].view(self._rew_shape), b["te"], b["tr"], b["final"].view(self._obs_shape)

    # ------------------------------------------------------------------ info on demand
    def get_info(self, env_index: int) -> dict:
        recs = self.engine.get_state_host().reshape(self.num_envs, self.num_cars, L.RECORD_WORDS)
        return I.env_info(recs[env_index])

    def close(self):
        if not self.closed:
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
            return venv.reset()[0]

        def step_async(self, actions):
            self._actions = actions

        def step_wait(self):
            obs, rew, te, tr, info = venv.step(self._actions)
            done = te | tr
            infos = [{} for _ in range(num_envs)]
            if info:
                for e in np.nonzero(done)[0]:
                    infos[e] = {"terminal_observation": info["final_observation"][e], "TimeLimit.truncated": bool(tr[e] and not te[e]),
                                "episode": {"r": float(info["episode"]["r"][e]), "l": int(info["episode"]["l"][e])}}
            return obs, rew, done, infos

        def close(self):
            venv.close()

        def get_attr(self, attr_name, indices=None):
            return [getattr(venv, attr_name)] * len(self._get_indices(indices))

        def set_attr(self, attr_name, value, indices=None):
            setattr(venv, attr_name, value)

        def env_method(self, method_name, *args, indices=None, **kwargs):
            return [getattr(venv, method_name)(*args, **kwargs)] * len(self._get_indices(indices))

        def env_is_wrapped(self, wrapper_class, indices=None):
            return [False] * len(self._get_indices(indices))

        def seed(self, seed=None):
            return [seed] * num_envs

    return _SB3()
