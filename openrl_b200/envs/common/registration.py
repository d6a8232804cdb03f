"""`make()` with the reference's signature (openrl/envs/common/registration.py:35-182).

Ids with a CUDA step function return a `DeviceVecEnv`; anything else is not part of the hot path
this package replaces and raises (the reference's CPU vec-envs keep serving those ids)."""
from typing import Callable, Optional

from .. import _kinds
from ..vec_env.device_venv import DeviceVecEnv


def make(id: str, env_num: int = 1, asynchronous: bool = False, add_monitor: bool = True, render_mode=None,
         make_custom_envs: Optional[Callable] = None, auto_reset: bool = True, device="cuda:0", **kwargs):
    if id not in _kinds.ENV_SPECS:
        raise NotImplementedError(
            f"env id {id!r} has no device-resident step function in openrl_b200 "
            f"(supported: {sorted(_kinds.ENV_SPECS)}); use the reference's make() for host envs")
    if not auto_reset:
        raise NotImplementedError("auto_reset=False is not supported by the device vec-env")
    return DeviceVecEnv(id, env_num, device=device, **kwargs)
