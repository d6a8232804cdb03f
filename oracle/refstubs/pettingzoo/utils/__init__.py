from . import env, wrappers  # noqa: F401
