from .device_venv import DeviceVecEnv  # noqa: F401
