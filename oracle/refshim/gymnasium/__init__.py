"""TEST INFRASTRUCTURE: the sliver of gymnasium 0.29 that /root/reference/src/base_env.py touches
(gym.Env, spaces.Box/Discrete/MultiDiscrete with contains()).  Only used by oracle/gen_golden.py."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def reset(self, seed=None, options=None):
        return None
