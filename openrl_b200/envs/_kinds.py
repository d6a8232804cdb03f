"""Registry of envs with a device step function (ids as in the reference's make())."""
import numpy as np

from .. import lib, spaces


def _cartpole_space():
    high = np.array([4.8, np.finfo(np.float32).max, 0.41887903, np.finfo(np.float32).max], np.float32)
    return spaces.Box(-high, high, dtype=np.float32)


def _gridworld_space():
    # 10x10 through make() (openrl/envs/gridworld/gridworld_env.py:14-18)
    return spaces.Box(low=np.array([0, 0, 0, 0]), high=np.array([9, 9, 9, 9]), dtype=np.int64)


def _mpe_space():
    # MultiAgentEnv.observation_space (multiagent_env.py:146-151): Dict{"policy": Box(18), "critic": Box(54)}
    return spaces.Dict({"policy": spaces.Box(-np.inf, np.inf, (18,), np.float32),
                        "critic": spaces.Box(-np.inf, np.inf, (54,), np.float32)})


def _gridworld2p_space():
    # (x0, y0, x1, y1) of the two players on the 10x10 grid (csrc/orl_selfplay.cu)
    return spaces.Box(low=np.array([0, 0, 0, 0]), high=np.array([9, 9, 9, 9]), dtype=np.int64)


ENV_SPECS = {
    "GridWorldSelfPlay": dict(kind=lib.ENV_GRIDWORLD_2P, agents=1, obs_dim=4, n_actions=5, i32_rows=8, observation_space=_gridworld2p_space),
    "simple_spread": dict(kind=lib.ENV_MPE_SPREAD, agents=3, obs_dim=18, critic_obs_dim=54, n_actions=5,
                          f64_rows=18, observation_space=_mpe_space),
    "CartPole-v1": dict(kind=lib.ENV_CARTPOLE, agents=1, obs_dim=4, n_actions=2, observation_space=_cartpole_space),
    "GridWorldEnv": dict(kind=lib.ENV_GRIDWORLD, agents=1, obs_dim=4, n_actions=5, observation_space=_gridworld_space),
}
