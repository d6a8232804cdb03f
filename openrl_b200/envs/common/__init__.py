from .registration import make  # noqa: F401
