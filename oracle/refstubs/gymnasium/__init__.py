"""Minimal stand-in for `gymnasium` (not installed in this image, no network).

TEST INFRASTRUCTURE ONLY: lets the *unmodified* reference (/root/reference) be imported
and executed in the build container so that golden vectors can be generated from it
(oracle/gen_golden.py).  Nothing in the product path imports this.
Only the attributes the reference's hot path touches are provided.
"""
from . import error, logger, spaces  # noqa: F401
from .spaces import Space  # noqa: F401
from .core import (  # noqa: F401
    ActionWrapper,
    Env,
    ObservationWrapper,
    RewardWrapper,
    Wrapper,
)
from . import core, utils, vector, wrappers, envs  # noqa: F401,E402
from .envs.registration import make, register  # noqa: F401,E402

__version__ = "0.29.1-stub"
