"""Registry with one entry: CartPole-v1, backed by the float64 restatement in
oracle/cartpole_ref.py (gymnasium's classic_control/cartpole.py is third-party and
absent from /root/reference)."""


class EnvSpec:
    def __init__(self, id, entry_point=None, max_episode_steps=None, kwargs=None):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.kwargs = kwargs or {}


registry = {}


def register(id, entry_point, max_episode_steps=None, **kwargs):
    registry[id] = EnvSpec(id, entry_point, max_episode_steps, kwargs.get("kwargs"))


def spec(id):
    return registry[id]


def make(id, render_mode=None, disable_env_checker=None, **kwargs):
    from ..wrappers import TimeLimit

    s = registry[id] if isinstance(id, str) else id
    env = s.entry_point(render_mode=render_mode, **{**s.kwargs, **kwargs})
    env.spec = s
    if s.max_episode_steps is not None:
        env = TimeLimit(env, s.max_episode_steps)
    return env


def _cartpole(render_mode=None, **kwargs):
    from cartpole_ref import CartPoleEnv

    return CartPoleEnv(render_mode=render_mode)


register("CartPole-v1", _cartpole, max_episode_steps=500)
