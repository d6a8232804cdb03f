from .ppo import PPOAlgorithm  # noqa: F401
