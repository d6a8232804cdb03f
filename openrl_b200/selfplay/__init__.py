from .opponent_pool import OpponentPool  # noqa: F401
