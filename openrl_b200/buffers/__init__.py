from .normal_buffer import NormalReplayBuffer  # noqa: F401
from .replay_data import ReplayData  # noqa: F401
