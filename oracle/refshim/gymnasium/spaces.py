import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            try:
                x = np.asarray(x, dtype=self.dtype)
            except (ValueError, TypeError):
                return False
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))


class Discrete:
    def __init__(self, n):
        self.n = int(n)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            return 0 <= int(x) < self.n
        if isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
            return 0 <= int(x) < self.n
        return False


class MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.nvec.shape and np.issubdtype(x.dtype, np.integer) and np.all(x >= 0) and np.all(x < self.nvec))
