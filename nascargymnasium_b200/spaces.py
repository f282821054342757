"""Action/observation spaces of the CarEnv surface (/root/reference/src/base_env.py:56-99).

Uses gymnasium.spaces when gymnasium is importable (so SB3/gymnasium wrappers see the real classes);
otherwise a minimal stand-in with the same constructor arguments, ``contains`` and ``sample``
(gymnasium is not installable in the offline build image)."""
from __future__ import annotations

import numpy as np

from . import constants as K

try:  # pragma: no cover - depends on the environment
    from gymnasium import spaces as _gs
    Box, Discrete, MultiDiscrete = _gs.Box, _gs.Discrete, _gs.MultiDiscrete
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self):
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            super().__init__()
            self.dtype = np.dtype(dtype)
            shape = tuple(shape) if shape is not None else np.shape(low)
            self.shape = shape
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), shape).copy()

        def contains(self, x) -> bool:
            if not isinstance(x, np.ndarray):
                try:
                    x = np.asarray(x, dtype=self.dtype)
                except (ValueError, TypeError):
                    return False
            return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n: int):
            super().__init__()
            self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)

        def contains(self, x) -> bool:
            if isinstance(x, (int, np.integer)):
                return 0 <= int(x) < self.n
            if isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
                return 0 <= int(x) < self.n
            return False

        def sample(self):
            return int(self._rng.integers(0, self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(_Space):
        def __init__(self, nvec):
            super().__init__()
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.shape, self.dtype = self.nvec.shape, np.dtype(np.int64)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.issubdtype(x.dtype, np.integer) and np.all(x >= 0) and np.all(x < self.nvec))

        def sample(self):
            return self._rng.integers(0, self.nvec).astype(self.dtype)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"


CAR_ACTION_LOW = np.array([-1.0, -1.0], dtype=np.float32)
CAR_ACTION_HIGH = np.array([1.0, 1.0], dtype=np.float32)
CAR_OBSERVATION_LOW = np.array(K.OBS_LOW, dtype=np.float32)
CAR_OBSERVATION_HIGH = np.array(K.OBS_HIGH, dtype=np.float32)


def make_spaces(discrete_action_space: bool, num_cars: int):
    """(action_space, observation_space) exactly as BaseEnv.__init__ builds them (base_env.py:56-99)."""
    if discrete_action_space:
        action = Discrete(5) if num_cars == 1 else MultiDiscrete([5] * num_cars)
    elif num_cars == 1:
        action = Box(low=CAR_ACTION_LOW, high=CAR_ACTION_HIGH, shape=(2,), dtype=np.float32)
    else:
        action = Box(low=np.tile(CAR_ACTION_LOW, (num_cars, 1)), high=np.tile(CAR_ACTION_HIGH, (num_cars, 1)), shape=(num_cars, 2),
                     dtype=np.float32)
    if num_cars == 1:
        obs = Box(low=CAR_OBSERVATION_LOW, high=CAR_OBSERVATION_HIGH, shape=(K.OBS_DIM,), dtype=np.float32)
    else:
        obs = Box(low=np.tile(CAR_OBSERVATION_LOW, (num_cars, 1)), high=np.tile(CAR_OBSERVATION_HIGH, (num_cars, 1)),
                  shape=(num_cars, K.OBS_DIM), dtype=np.float32)
    return action, obs
