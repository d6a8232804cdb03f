import numpy as np

from .space import Space


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self, mask=None):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(0 <= x) and np.all(x < self.nvec))


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n
        shape = (n,) if np.isscalar(n) else tuple(n)
        super().__init__(shape, np.int8, seed)

    def sample(self, mask=None):
        return self.np_random.integers(0, 2, size=self.shape, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(p) for s, p in zip(self.spaces, x))


class Graph(Space):
    pass


class Sequence(Space):
    pass


class Text(Space):
    pass


class GraphInstance:
    pass
