"""evaluate_policy(agent, env, n_eval_episodes, deterministic, ...) on the device path
(reference: openrl/utils/evaluation.py:13-165).

Same contract: the episodes are divided over the envs of the vector env up front (env i plays
(n_eval_episodes + i) // n_envs episodes, so no env's faster episodes bias the estimate), the agent's env is
swapped for the evaluation env and restored afterwards, recurrent agents get `episode_starts` so that their hidden
state is reset with the episodes, an episode ends when ALL agents of the env are done and its return is the running
per-agent sum.  The vector envs here auto-reset and report whole-episode statistics themselves; rewards are
accumulated from `env.step` as in the reference's non-Monitor branch."""
import numpy as np


def evaluate_policy(agent, env, n_eval_episodes=10, deterministic=True, render=False, callback=None, reward_threshold=None,
                    return_episode_rewards=False, warn=True):
    n_envs = env.parallel_env_num
    targets = np.array([(n_eval_episodes + i) // n_envs for i in range(n_envs)], dtype=int)
    counts = np.zeros(n_envs, dtype=int)
    current_rewards = np.zeros([n_envs, env.agent_num])
    current_lengths = np.zeros(n_envs, dtype=int)
    episode_rewards, episode_lengths = [], []

    train_env = agent.get_env()
    agent.set_env(env)
    try:
        reset = env.reset()
        observations = reset[0] if isinstance(reset, tuple) else reset
        episode_starts = np.ones((n_envs,), dtype=bool)
        while (counts < targets).any():
            starts = episode_starts if episode_starts.any() else None
            actions, _ = agent.act(observations, deterministic=deterministic, episode_starts=starts)
            observations, rewards, dones, infos = env.step(actions)
            current_rewards += np.squeeze(np.asarray(rewards), axis=-1)
            current_lengths += 1
            for i in range(n_envs):
                if counts[i] >= targets[i]:
                    continue
                reward, info = rewards[i], (infos[i] if infos is not None else {})  # noqa: F841  (callback locals)
                done = bool(np.all(dones[i]))
                episode_starts[i] = done
                if callback is not None:
                    callback(locals(), globals())
                if done:
                    episode_rewards.append(current_rewards[i].copy())
                    episode_lengths.append(int(current_lengths[i]))
                    counts[i] += 1
                    current_rewards[i] = 0
                    current_lengths[i] = 0
    finally:
        agent.set_env(train_env)
    mean_reward, std_reward = np.mean(episode_rewards), np.std(episode_rewards)
    if reward_threshold is not None:
        assert mean_reward > reward_threshold, f"Mean reward below threshold: {mean_reward:.2f} < {reward_threshold:.2f}"
    if return_episode_rewards:
        return episode_rewards, episode_lengths
    return mean_reward, std_reward
