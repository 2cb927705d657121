import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = dtype
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(tuple(shape) if shape is not None else None, np.dtype(dtype))
        self.low, self.high = low, high

    def sample(self):
        return self._rng.uniform(self.low, self.high, size=self.shape).astype(self.dtype)


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n, self.start = int(n), int(start)

    def sample(self):
        return int(self._rng.integers(self.n)) + self.start


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, np.dtype(dtype))

    def sample(self):
        return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        super().__init__(None, None)
        self.spaces = dict(spaces or {}, **kw)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}
