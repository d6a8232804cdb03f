from .ppo_net import PPONet  # noqa: F401
