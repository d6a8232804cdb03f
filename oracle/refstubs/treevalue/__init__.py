"""dict-like TreeValue stand-in (reference uses it in openrl/buffers/utils/obs_data.py only)."""


class TreeValue:
    def __init__(self, data):
        object.__setattr__(self, "_d", dict(data._d if isinstance(data, TreeValue) else data))

    def keys(self):
        return self._d.keys()

    def values(self):
        return self._d.values()

    def items(self):
        return self._d.items()

    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k] = v

    def __contains__(self, k):
        return k in self._d

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_d")[k]
        except KeyError:
            raise AttributeError(k)

    def __len__(self):
        return len(self._d)


def reduce_(tree, fn):
    return fn(**{k: v for k, v in tree.items()})
