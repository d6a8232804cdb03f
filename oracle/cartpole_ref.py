"""CPU restatement of gymnasium 0.29 `CartPole-v1` (classic_control/cartpole.py).

TEST INFRASTRUCTURE (oracle).  gymnasium is a third-party dependency of the reference
(`setup.py:28`, "gymnasium>=0.29", unpinned, not vendored) and is absent from this image,
so the dynamics are restated here from the published algorithm (Barto, Sutton & Anderson
1983 cart-pole, explicit Euler, tau = 0.02) and registered in the stand-in `gymnasium`
registry so that the *unmodified* reference trains on it.  Parity at this boundary is
UNPINNED by the reference's own tests (only `tests/test_examples/test_train_cartpole.py:53`
"return >= 450" exists); reference call sites: `openrl/envs/common/registration.py:91`,
`openrl/envs/gymnasium/__init__.py:46-53`, `openrl/envs/common/build_envs.py:40-45`.

State is float64 numpy, observation is cast to float32, reset draws
U(-0.05, 0.05)^4 from the env's PCG64 `np_random` (seeded by reset(seed=...)).
"""
import math

import numpy as np

try:  # the stand-in gymnasium (oracle/refstubs) when generating goldens
    import gymnasium as gym
    from gymnasium import spaces

    _Base = gym.Env
except Exception:  # pragma: no cover
    gym = None
    spaces = None
    _Base = object

GRAVITY = 9.8
MASSCART = 1.0
MASSPOLE = 0.1
TOTAL_MASS = MASSPOLE + MASSCART
LENGTH = 0.5  # half the pole's length
POLEMASS_LENGTH = MASSPOLE * LENGTH
FORCE_MAG = 10.0
TAU = 0.02
THETA_THRESHOLD_RADIANS = 12 * 2 * math.pi / 360
X_THRESHOLD = 2.4
MAX_EPISODE_STEPS = 500


def cartpole_step_f64(state, action):
    """One Euler step on a float64 state (x, x_dot, theta, theta_dot).

    Returns (new_state, terminated)."""
    x, x_dot, theta, theta_dot = state
    force = FORCE_MAG if action == 1 else -FORCE_MAG
    costheta = math.cos(theta)
    sintheta = math.sin(theta)
    temp = (force + POLEMASS_LENGTH * theta_dot**2 * sintheta) / TOTAL_MASS
    thetaacc = (GRAVITY * sintheta - costheta * temp) / (
        LENGTH * (4.0 / 3.0 - MASSPOLE * costheta**2 / TOTAL_MASS)
    )
    xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / TOTAL_MASS
    x = x + TAU * x_dot
    x_dot = x_dot + TAU * xacc
    theta = theta + TAU * theta_dot
    theta_dot = theta_dot + TAU * thetaacc
    terminated = bool(
        x < -X_THRESHOLD
        or x > X_THRESHOLD
        or theta < -THETA_THRESHOLD_RADIANS
        or theta > THETA_THRESHOLD_RADIANS
    )
    return (x, x_dot, theta, theta_dot), terminated


class CartPoleEnv(_Base):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 50}

    def __init__(self, render_mode=None):
        high = np.array(
            [
                X_THRESHOLD * 2,
                np.finfo(np.float32).max,
                THETA_THRESHOLD_RADIANS * 2,
                np.finfo(np.float32).max,
            ],
            dtype=np.float32,
        )
        self.action_space = spaces.Discrete(2)
        self.observation_space = spaces.Box(-high, high, dtype=np.float32)
        self.render_mode = render_mode
        self.state = None
        self.steps_beyond_terminated = None

    def step(self, action):
        assert self.state is not None, "Call reset before using step method."
        self.state, terminated = cartpole_step_f64(tuple(self.state), int(action))
        if not terminated:
            reward = 1.0
        elif self.steps_beyond_terminated is None:
            self.steps_beyond_terminated = 0
            reward = 1.0
        else:
            self.steps_beyond_terminated += 1
            reward = 0.0
        return np.array(self.state, dtype=np.float32), reward, terminated, False, {}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        self.state = self.np_random.uniform(low=-0.05, high=0.05, size=(4,))
        self.steps_beyond_terminated = None
        return np.array(self.state, dtype=np.float32), {}

    def close(self):
        pass
