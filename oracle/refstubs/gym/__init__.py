"""`gym` alias of the gymnasium stand-in (test infrastructure only)."""
import sys

import gymnasium
from gymnasium import *  # noqa: F401,F403
from gymnasium import Env, Wrapper, error, spaces, envs, core, utils, wrappers  # noqa: F401

for _name in ("spaces", "error", "envs", "core", "utils", "wrappers", "spaces.dict", "spaces.box", "spaces.discrete"):
    sys.modules["gym." + _name] = sys.modules["gymnasium." + _name]
