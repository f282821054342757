"""``CarEnv``: the reference's Gymnasium surface (/root/reference/src/car_env.py:79-803, src/base_env.py:42-136)
over the CUDA engine -- same constructor kwargs, spaces, return shapes, error behaviour and info keys, one env
of 1..10 cars per instance.  For throughput use :class:`nascargymnasium_b200.vector_env.NascarVectorEnv`;
this class exists so single-env callers (game/*.py, learn/genetic_trainer.py, demo/random_demo.py, SB3's
DummyVecEnv/Monitor) can switch imports and keep working.  No CPU fallback: construction needs a CUDA device."""
from __future__ import annotations

import glob
import os
import random
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import constants as K
from . import info as I
from . import layout as L
from . import spaces as S
from . import track as T
from .engine import Engine

try:  # pragma: no cover
    import gymnasium as _gym
    _Base = _gym.Env
except Exception:
    _Base = object


class CarEnv(_Base):
    metadata = {"render_modes": ["human"], "render_fps": 60}

    def __init__(self, render_mode: Optional[str] = None, track_file: Optional[str] = None,
                 start_position: Optional[Tuple[float, float]] = None, start_angle: float = 0.0, reset_on_lap: bool = False,
                 discrete_action_space: bool = False, num_cars: int = 1, car_names: Optional[list] = None, device: int = 0,
                 car_contacts: bool = False, start_grid: Tuple[float, float] = (8.0, 3.0)):
        # car_contacts (NOT a reference kwarg, default off = the reference's behaviour: every car alone in its own b2World,
        # car_env.py:389-394): the cars share one world, start on a two-wide grid (rows start_grid[0] m apart, lanes
        # start_grid[1] m apart) and collide with each other (NcgConfig.car_contacts)
        self.car_contacts, self.start_grid = bool(car_contacts), (float(start_grid[0]), float(start_grid[1]))
        if num_cars < 1 or num_cars > K.MAX_CARS:
            raise ValueError(f"Number of cars must be between 1 and {K.MAX_CARS}")
        if car_names is None:
            self.car_names = [f"Car {i}" for i in range(num_cars)]
        else:
            if len(car_names) != num_cars:
                raise ValueError(f"Number of car names ({len(car_names)}) must match number of cars ({num_cars})")
            self.car_names = list(car_names)
        # render_mode="human": the window is the reference's own pygame Renderer, fed from the engine's records through
        # render_bridge.frame_kwargs; it is created at the first render() and needs the reference tree + pygame importable
        self._renderer = None
        self._last_rewards, self._last_actions = None, None
        self.render_mode = render_mode
        self.discrete_action_space = discrete_action_space
        self.num_cars = num_cars
        self.reset_on_lap = reset_on_lap
        # car_env.py:114-115; (0, 0) is replaced by the GRID/STARTLINE segment's start (:236-241), which is the origin on
        # every track the generator can produce (Track.add_segment starts at (0, 0))
        self.start_position = tuple(float(v) for v in start_position) if start_position else (0.0, 0.0)
        self.start_angle = float(start_angle)
        self.followed_car_index = 0
        self.action_space, self.observation_space = S.make_spaces(discrete_action_space, num_cars)
        self.track_file = track_file
        self._is_random_track_mode = track_file is None
        if track_file is not None:
            T.load_track(track_file)            # FileNotFoundError / ValueError like TrackLoader.load_track
            self._tracks = [track_file]
        else:
            self._tracks = self._discover_available_tracks()
        self._device = device
        self._engine: Optional[Engine] = None
        self._track_index = 0
        self.disabled_cars: set = set()
        self.cumulative_collision_impacts: Dict[int, float] = {}
        self.termination_reason = None
        self.simulation_time = 0.0
        self._was_reset = False
        self._cumulative_rewards = [0.0] * num_cars

    # ------------------------------------------------------------------ tracks (car_env.py:243-303)
    def _discover_available_tracks(self) -> List[str]:
        here = os.path.join(os.getcwd(), "tracks")
        files = sorted(glob.glob(os.path.join(here, "*.track")))
        return files if files else [f"tracks/{n}.track" for n in T.BUILTIN_TRACK_NAMES]

    def switch_to_random(self):
        self.track_file = None
        self._is_random_track_mode = True
        self._tracks = self._discover_available_tracks()
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def seed(self, seed_value: int = None) -> list:
        if seed_value is None:
            seed_value = random.randint(0, 2 ** 32 - 1)
        random.seed(seed_value)
        np.random.seed(seed_value)
        return [seed_value]

    def _ensure_engine(self):
        if self._engine is None:
            self._engine = Engine(1, self.num_cars, tracks=self._tracks, discrete=self.discrete_action_space,
                                  reset_on_lap=self.reset_on_lap, auto_reset=False, device=self._device, track_info=True,
                                  start_position=self.start_position, start_angle=self.start_angle,
                                  car_contacts=self.car_contacts, grid=self.start_grid)

    # ------------------------------------------------------------------ gym API
    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
        self.seed(seed)
        self._ensure_engine()
        fresh = not self._was_reset
        if self._is_random_track_mode and len(self._tracks) > 1:
            prev = self._track_index if self._was_reset else None
            choices = [i for i in range(len(self._tracks)) if i != prev]
            random.seed(os.getpid() + int(time.time() * 1000) % 1000)       # car_env.py:281-282
            self._track_index = random.choice(choices)
            fresh = True                                                    # track changed => new physics worlds
            self.track_file = self._tracks[self._track_index]
        elif self._is_random_track_mode:
            self.track_file = self._tracks[0]
        obs = self._engine.reset_host(track_id=np.array([self._track_index], dtype=np.int32), fresh=fresh)
        self._was_reset = True
        self.disabled_cars = set()
        self.cumulative_collision_impacts = {i: 0.0 for i in range(self.num_cars)}
        self._cumulative_rewards = [0.0] * self.num_cars
        self.termination_reason = None
        self.simulation_time = 0.0
        info = self._info(self._engine.get_state_host())
        return (obs[0], info) if self.num_cars == 1 else (obs.reshape(self.num_cars, K.OBS_DIM), info)

    def step(self, action):
        assert self.action_space.contains(action), f"Invalid action {action}"
        if not self._was_reset:
            raise RuntimeError("Environment not properly initialized. Call reset() first.")
        if self.discrete_action_space:
            a = np.asarray(action, dtype=np.int32).reshape(self.num_cars)
        else:
            a = np.asarray(action, dtype=np.float32).reshape(self.num_cars, 2)
        obs, rew, te, tr, _ = self._engine.step_host(a)
        self._last_rewards, self._last_actions = rew, a
        recs = self._engine.get_state_host()
        terminated, truncated = bool(te[0]), bool(tr[0])
        self._update_mirrors(recs, rew, terminated, truncated)
        info = self._info(recs)
        if self.num_cars == 1:
            return obs[0], rew[0], terminated, truncated, info
        return obs.reshape(self.num_cars, K.OBS_DIM), rew, terminated, truncated, info

    def _update_mirrors(self, recs, rew, terminated, truncated):
        fl = recs.view(np.uint32)[:, L.R["NCG_R_FLAGS"]]
        self.disabled_cars = {i for i in range(self.num_cars) if fl[i] & L.F["NCG_F_DISABLED"]}
        self.cumulative_collision_impacts = {i: float(recs[i, L.R["NCG_R_CUM_IMPACT"]]) for i in range(self.num_cars)}
        prev = list(self._cumulative_rewards)
        self._cumulative_rewards = [np.float32(recs[i, L.R["NCG_R_CUM_REWARD"]]) for i in range(self.num_cars)]
        step = int(recs.view(np.uint32)[0, L.R["NCG_R_STEP"]])
        self.simulation_time = I.sim_time(step)
        # termination_reason (car_env.py:1115-1158); a lap reset terminates without a reason
        reason = None
        active = [i for i in range(self.num_cars) if i not in self.disabled_cars]
        if len(self.disabled_cars) >= self.num_cars:
            reason = I.TERMINATION_REASONS[1]
        elif active and all(prev[i] < K.TERMINATION_MIN_REWARD for i in active):
            reason = I.TERMINATION_REASONS[2]
        elif self.reset_on_lap and step >= K.TERMINATION_STEPS:
            reason = I.TERMINATION_REASONS[3]
        elif step >= K.TRUNCATION_STEPS:
            reason = I.TERMINATION_REASONS[4]
        self.termination_reason = reason

    def _info(self, recs) -> Dict[str, Any]:
        info = I.env_info(recs, self.termination_reason, self.followed_car_index, hist=self._engine.velocity_history_host())
        n_bodies = 1 + self._engine.tables[self._track_index].n_walls
        for p in info["physics"]:
            p["bodies_in_world"] = n_bodies
        return info

    def render_state(self) -> Dict[str, Any]:
        """What CarEnv.render passes to Renderer.render_frame (car_env.py:1387-1401), from the engine's records."""
        from . import render_bridge as RB
        if self._engine is None or not self._was_reset:
            raise RuntimeError("Environment not properly initialized. Call reset() first.")
        tab = self._engine.tables[self._track_index]
        return RB.frame_kwargs(self._engine.get_state_host(), tab.seg64, tab.track.total_length, self.car_names, self.followed_car_index,
                               self._last_rewards, self._last_actions, self.reset_on_lap, self.track_file)

    def render(self):
        if self.render_mode != "human":
            return None
        from . import render_bridge as RB
        if self._renderer is None:
            try:
                self._renderer = RB.make_reference_renderer(self.track_file, self.metadata["render_fps"])
            except ImportError as e:
                raise NotImplementedError("render_mode='human' draws with the reference's own pygame Renderer (src/renderer.py): put the "
                                          f"NascarGymnasium tree on sys.path and install pygame ({e}); render_state() gives the frame data") from e
        if self._was_reset:
            self._renderer.render_frame(**self.render_state())

    def check_quit_requested(self) -> bool:
        return False

    def close(self) -> None:
        if self._renderer is not None:
            self._renderer.close()
            self._renderer = None
        if self._engine is not None:
            self._engine.close()
            self._engine = None
