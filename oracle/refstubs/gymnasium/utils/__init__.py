from . import seeding  # noqa: F401
from . import step_api_compatibility  # noqa: F401


class RecordConstructorArgs:
    def __init__(self, *a, **k):
        pass


class EzPickle:
    def __init__(self, *a, **k):
        pass
