from collections import OrderedDict

from .space import Space


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **spaces_kwargs):
        if spaces is None:
            spaces = {}
        if isinstance(spaces, (list, tuple)):
            spaces = OrderedDict(spaces)
        spaces = dict(spaces)
        spaces.update(spaces_kwargs)
        self.spaces = spaces
        super().__init__(None, None, seed)

    def __getitem__(self, key):
        return self.spaces[key]

    def __setitem__(self, key, value):
        self.spaces[key] = value

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def sample(self, mask=None):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def seed(self, seed=None):
        out = []
        for i, s in enumerate(self.spaces.values()):
            out += s.seed(None if seed is None else seed + i)
        return out

    def contains(self, x):
        return isinstance(x, dict) and all(k in x and self.spaces[k].contains(x[k]) for k in self.spaces)

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {s}" for k, s in self.spaces.items()) + ")"

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces
