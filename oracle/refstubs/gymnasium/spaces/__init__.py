from .space import Space
from .box import Box
from .discrete import Discrete
from .dict import Dict
from .misc import Graph, GraphInstance, MultiBinary, MultiDiscrete, Sequence, Text, Tuple
from .utils import flatdim, flatten, flatten_space, unflatten

__all__ = [
    "Space", "Box", "Discrete", "Dict", "Graph", "GraphInstance", "MultiBinary", "MultiDiscrete",
    "Sequence", "Text", "Tuple", "flatdim", "flatten", "flatten_space", "unflatten",
]
