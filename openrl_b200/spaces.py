"""Minimal observation / action spaces (gymnasium is not a dependency of this package).

Only what the hot path reads: `.shape`, `.n`, `.dtype`, `.low/.high`, `.sample()`,
`__class__.__name__` in ("Box", "Discrete", "Dict") — the reference dispatches on the class
name (openrl/modules/networks/utils/act.py:14-44, openrl/buffers/utils/util.py).
"""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._rng = np.random.default_rng()

    @property
    def shape(self):
        return self._shape

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape if not np.isscalar(low) else (1,)
        shape = tuple(int(s) for s in shape)
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), shape).copy()

    def sample(self, mask=None):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def __eq__(self, o):
        return isinstance(o, Box) and self.shape == o.shape and np.allclose(self.low, o.low) and np.allclose(self.high, o.high)


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)

    def sample(self, mask=None):
        if mask is not None:
            valid = np.flatnonzero(np.asarray(mask) == 1)
            if len(valid):
                return self.start + int(self._rng.choice(valid))
            return self.start
        return self.start + int(self._rng.integers(self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, o):
        return isinstance(o, Discrete) and self.n == o.n and self.start == o.start


class Dict(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {v}" for k, v in self.spaces.items()) + ")"

    def __eq__(self, o):
        return isinstance(o, Dict) and self.spaces == o.spaces
