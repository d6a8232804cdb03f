import numpy as np

from .space import Space


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = np.int64(n)
        self.start = np.int64(start)
        super().__init__((), np.int64, seed)

    @property
    def is_np_flattenable(self):
        return True

    def sample(self, mask=None):
        if mask is not None:
            valid = mask == 1
            if np.any(valid):
                return self.start + self.np_random.choice(np.where(valid)[0])
            return self.start
        return self.start + self.np_random.integers(self.n)

    def contains(self, x):
        try:
            xi = int(x)
        except Exception:
            return False
        return bool(self.start <= xi < self.start + self.n)

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start
