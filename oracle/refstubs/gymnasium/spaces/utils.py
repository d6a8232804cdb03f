import numpy as np

from .box import Box
from .dict import Dict
from .discrete import Discrete
from .misc import MultiBinary, MultiDiscrete, Tuple


def flatdim(space):
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    if isinstance(space, MultiDiscrete):
        return int(np.sum(space.nvec))
    if isinstance(space, MultiBinary):
        return int(np.prod(space.shape))
    if isinstance(space, (Tuple,)):
        return sum(flatdim(s) for s in space.spaces)
    if isinstance(space, Dict):
        return sum(flatdim(s) for s in space.spaces.values())
    raise NotImplementedError(space)


def flatten(space, x):
    if isinstance(space, Box):
        return np.asarray(x, dtype=space.dtype).flatten()
    if isinstance(space, Discrete):
        onehot = np.zeros(space.n, dtype=np.int64)
        onehot[int(x) - int(space.start)] = 1
        return onehot
    raise NotImplementedError(space)


def unflatten(space, x):
    if isinstance(space, Box):
        return np.asarray(x, dtype=space.dtype).reshape(space.shape)
    if isinstance(space, Discrete):
        return int(space.start + np.nonzero(x)[0][0])
    raise NotImplementedError(space)


def flatten_space(space):
    if isinstance(space, Box):
        return Box(space.low.flatten(), space.high.flatten(), dtype=space.dtype)
    if isinstance(space, Discrete):
        return Box(low=0, high=1, shape=(int(space.n),), dtype=np.int64)
    raise NotImplementedError(space)
