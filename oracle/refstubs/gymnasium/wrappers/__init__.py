from ..core import Wrapper
from ..utils.step_api_compatibility import step_api_compatibility


class StepAPICompatibility(Wrapper):
    def __init__(self, env, output_truncation_bool=True):
        super().__init__(env)
        self.is_vector_env = False
        self.output_truncation_bool = output_truncation_bool

    def step(self, action):
        step_returns = self.env.step(action)
        return step_api_compatibility(step_returns, self.output_truncation_bool, self.is_vector_env)


class AutoResetWrapper(Wrapper):
    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        if terminated or truncated:
            new_obs, new_info = self.env.reset()
            new_info["final_observation"] = obs
            new_info["final_info"] = info
            obs, info = new_obs, new_info
        return obs, reward, terminated, truncated, info


class EnvCompatibility(Wrapper):
    pass


class TimeLimit(Wrapper):
    """gymnasium 0.29 wrappers/time_limit.py semantics."""

    def __init__(self, env, max_episode_steps=None):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return observation, reward, terminated, truncated, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)
