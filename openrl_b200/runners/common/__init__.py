from .ppo_agent import PPOAgent  # noqa: F401
