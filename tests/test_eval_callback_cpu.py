"""Host logic of evaluate_policy / EvalCallback (reference: openrl/utils/evaluation.py:13-165,
utils/callbacks/eval_callback.py:53-284) with a scripted vector env and agent — no GPU involved."""
import os

import numpy as np

from openrl_b200.utils.callbacks import BaseCallback, CallbackList, EvalCallback
from openrl_b200.utils.evaluation import evaluate_policy


class ScriptedVecEnv:
    """Env i gives reward (i + 1) per step to each of A agents and ends every `lengths[i]` steps (auto-reset)."""

    def __init__(self, lengths, agents=2):
        self.lengths = np.asarray(lengths)
        self.parallel_env_num, self.agent_num = len(lengths), agents
        self.t = np.zeros(len(lengths), dtype=int)
        self.resets = 0
        self.closed = False

    def reset(self, seed=None, options=None):
        self.t[:] = 0
        self.resets += 1
        return np.zeros((self.parallel_env_num, self.agent_num, 3), np.float32), [{} for _ in self.lengths]

    def step(self, actions):
        assert actions.shape == (self.parallel_env_num, self.agent_num, 1)
        self.t += 1
        done = self.t >= self.lengths
        self.t[done] = 0
        rew = np.repeat((np.arange(self.parallel_env_num) + 1.0)[:, None, None], self.agent_num, axis=1)
        dones = np.repeat(done[:, None], self.agent_num, axis=1)
        return np.zeros((self.parallel_env_num, self.agent_num, 3), np.float32), rew, dones, [{} for _ in self.lengths]

    def close(self):
        self.closed = True


class ScriptedAgent:
    def __init__(self, env):
        self._env, self.env_num = env, env.parallel_env_num
        self.num_time_steps = 0
        self.starts, self.saved, self.logged = [], [], []
        self.logger = self

    def get_env(self):
        return self._env

    def set_env(self, env):
        self._env, self.env_num = env, env.parallel_env_num

    def act(self, obs, deterministic=True, episode_starts=None):
        self.starts.append(None if episode_starts is None else episode_starts.copy())
        return np.zeros((obs.shape[0], obs.shape[1], 1), np.int64), None

    def save(self, path):
        self.saved.append(path)

    def log_info(self, info, step):
        self.logged.append((dict(info), step))


def test_evaluate_policy_divides_episodes_and_restores_env():
    train_env, eval_env = ScriptedVecEnv([5, 5]), ScriptedVecEnv([2, 3, 4])
    agent = ScriptedAgent(train_env)
    rewards, lengths = evaluate_policy(agent, eval_env, n_eval_episodes=7, return_episode_rewards=True)
    # env i plays (7 + i) // 3 episodes: 2, 2, 3
    assert sorted(lengths) == sorted([2, 2, 3, 3, 4, 4, 4])
    by_len = {2: 1.0, 3: 2.0, 4: 3.0}
    for r, l in zip(rewards, lengths):
        np.testing.assert_array_equal(r, np.full(2, by_len[l] * l))     # per-agent running sums
    assert agent.get_env() is train_env and agent.env_num == 2            # training env restored
    assert agent.starts[0].all() and agent.starts[1] is None              # hidden-state resets only at episode starts
    assert any(s is not None and s.tolist() == [True, False, False] for s in agent.starts[2:])
    mean, std = evaluate_policy(agent, eval_env, n_eval_episodes=3)
    assert np.isclose(mean, np.mean([2.0, 6.0, 12.0])) and std > 0
    try:
        evaluate_policy(agent, eval_env, n_eval_episodes=3, reward_threshold=100.0)
        raise SystemExit("threshold not enforced")
    except AssertionError:
        pass
    assert agent.get_env() is train_env


def test_eval_callback_triggers_at_rollout_boundaries_saves_best_and_can_stop(tmp_path):
    class StopOnBest(BaseCallback):
        def _on_step(self):
            return self.parent.best_mean_reward < 5.0   # stop once the best mean reward reaches 5

    train_env, eval_env = ScriptedVecEnv([5, 5, 5, 5]), ScriptedVecEnv([2])
    agent = ScriptedAgent(train_env)
    cb = EvalCallback(eval_env, callbacks_on_new_best=StopOnBest(), n_eval_episodes=2, eval_freq=10, log_path=str(tmp_path),
                      best_model_save_path=str(tmp_path / "best"), verbose=0)
    assert cb.needs_per_step is False                # the driver may keep one-launch rollouts
    cbs = CallbackList([cb])
    cbs.init_callback(agent)
    cbs.on_training_start({}, {})
    conts = []
    for it in range(4):                              # rollouts of 8 vector steps on 4 envs
        agent.num_time_steps += 8 * 4
        conts.append(cbs.on_rollout_end())
    # vec steps 8, 16, 24, 32 -> evaluations due after crossing 10, 20, 30
    assert [s for _, s in agent.logged] == [64, 96, 128]
    assert agent.logged[0][0]["Eval/episode_reward"] == 2.0 and agent.logged[0][0]["Eval/episode_length"] == 2.0
    assert len(agent.saved) == 1 and agent.saved[0].endswith(os.path.join("best", "best_model"))   # only the first eval is a new best
    assert os.path.exists(tmp_path / "best" / "best_model_info.txt")
    z = np.load(str(tmp_path / "evaluations.npz"))
    assert z["timesteps"].tolist() == [64, 96, 128] and z["results"].shape[0] == 3
    assert conts == [True, True, True, True]         # best (2.0) never reaches the stop threshold
    eval_env.lengths[:] = 6                          # better policy: episodes of 6 steps -> reward 6 >= 5
    agent.num_time_steps += 8 * 4
    assert cbs.on_rollout_end() is False
    cbs.on_training_end()
    assert eval_env.closed


def test_checkpoint_and_stop_callbacks(tmp_path):
    from openrl_b200.utils.callbacks import (CheckpointCallback, StopTrainingOnMaxEpisodes, StopTrainingOnNoModelImprovement,
                                             StopTrainingOnRewardThreshold)

    train_env, eval_env = ScriptedVecEnv([5, 5, 5, 5]), ScriptedVecEnv([3])
    agent = ScriptedAgent(train_env)
    ck = CheckpointCallback(save_freq=16, save_path=str(tmp_path / "ckpt"), name_prefix="m")
    ev = EvalCallback(eval_env, callbacks_on_new_best=StopTrainingOnRewardThreshold(100.0),
                      callbacks_after_eval=StopTrainingOnNoModelImprovement(max_no_improvement_evals=1, min_evals=0),
                      n_eval_episodes=1, eval_freq=8, verbose=0, close_env_at_end=False)
    cbs = CallbackList([ck, ev])
    assert cbs.needs_per_step is False
    cbs.init_callback(agent)
    cbs.on_training_start({}, {})
    conts = []
    for _ in range(4):                               # rollouts of 8 vector steps on 4 envs
        agent.num_time_steps += 32
        conts.append(cbs.on_rollout_end())
    # evaluations: best = 3.0 at the first, then no improvement twice -> stop at the third evaluation
    assert conts[:3] == [True, True, False]
    assert [os.path.basename(p) for p in agent.saved if "ckpt" in p] == ["m_64_steps", "m_128_steps"]   # vec steps 16, 32
    assert os.path.isdir(tmp_path / "ckpt")
    # episode counter on per-step locals: 2 episodes per env over 4 envs
    mx = StopTrainingOnMaxEpisodes(max_episodes=2)
    assert mx.needs_per_step is True
    mx.init_callback(agent)
    mx.on_training_start({}, {})
    env = ScriptedVecEnv([2, 2, 2, 2], agents=1)
    env.reset()
    res = []
    for _ in range(6):
        _, _, dones, _ = env.step(np.zeros((4, 1, 1), np.int64))
        mx.update_locals({"dones": dones})
        res.append(mx.on_step())
    assert res == [True, True, True, False, False, False]   # 8 episodes reached at the 4th step
