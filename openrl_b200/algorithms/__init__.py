from .a2c import A2CAlgorithm  # noqa: F401
from .ppo import PPOAlgorithm  # noqa: F401
