from .device_venv import DeviceVecEnv  # noqa: F401
from .host_venv import HostVecEnv  # noqa: F401
from .host_sync import SyncHostVecEnv  # noqa: F401
