from .registration import make, register, registry, spec  # noqa: F401
from . import registration  # noqa: F401
