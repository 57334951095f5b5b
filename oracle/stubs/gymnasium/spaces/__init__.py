"""Discrete / MultiDiscrete with the gymnasium>=0.26 constructor signature."""
import numpy as np

Space = object


def _bind(space_cls):
    global Space, Discrete, MultiDiscrete
    Space = space_cls

    class Discrete(space_cls):
        def __init__(self, n, seed=None, start=0):
            self.n = int(n)
            self.start = int(start)
            super().__init__((), np.int64, seed)

        def sample(self, mask=None):
            from gymnasium.utils import seeding
            if self._np_random is None:
                self._np_random = seeding.np_random()[0]
            if mask is not None:
                valid = np.where(np.asarray(mask) == 1)[0]
                return int(self.start + self._np_random.choice(valid))
            return int(self.start + self._np_random.integers(self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

    class MultiDiscrete(space_cls):
        def __init__(self, nvec, dtype=np.int64, seed=None, start=None):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.start = np.zeros_like(self.nvec) if start is None else np.asarray(start, dtype=np.int64)
            super().__init__(self.nvec.shape, dtype, seed)

        def sample(self, mask=None):
            from gymnasium.utils import seeding
            if self._np_random is None:
                self._np_random = seeding.np_random()[0]
            return (self._np_random.random(self.nvec.shape) * self.nvec).astype(np.int64) + self.start

        def contains(self, x):
            x = np.asarray(x)
            return bool(np.all(x >= self.start) and np.all(x < self.start + self.nvec))

    globals()["Discrete"] = Discrete
    globals()["MultiDiscrete"] = MultiDiscrete
