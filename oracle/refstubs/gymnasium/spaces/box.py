import numpy as np

from .space import Space


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            if np.isscalar(low) and np.isscalar(high):
                shape = (1,)
            else:
                shape = np.asarray(low if not np.isscalar(low) else high).shape
        shape = tuple(int(s) for s in shape)
        self.low = np.full(shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low).astype(dtype).reshape(shape)
        self.high = np.full(shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high).astype(dtype).reshape(shape)
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        super().__init__(shape, dtype, seed)

    @property
    def is_np_flattenable(self):
        return True

    def sample(self, mask=None):
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp_bounded = ~self.bounded_below & self.bounded_above
        low_bounded = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
        sample[low_bounded] = self.np_random.exponential(size=low_bounded[low_bounded].shape) + self.low[low_bounded]
        sample[upp_bounded] = -self.np_random.exponential(size=upp_bounded[upp_bounded].shape) + high[upp_bounded]
        sample[bounded] = self.np_random.uniform(low=self.low[bounded], high=high[bounded], size=bounded[bounded].shape)
        if self.dtype.kind in ["i", "u", "b"]:
            sample = np.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return (
            isinstance(other, Box)
            and self.shape == other.shape
            and np.allclose(self.low, other.low)
            and np.allclose(self.high, other.high)
        )
