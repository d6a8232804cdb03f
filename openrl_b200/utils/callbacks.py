"""Callback hook points of the training loop (reference: openrl/utils/callbacks/callbacks.py:14-172;
call sites ppo_agent.py:128-132, onpolicy_driver.py:155,174-178,196).  The same duck type is
accepted: on_training_start(locals, globals), on_rollout_start(), update_locals(locals),
on_step() -> bool, on_rollout_end(), on_training_end()."""


class BaseCallback:
    def __init__(self, verbose=0):
        self.agent = None
        self.n_calls = 0
        self.num_time_steps = 0
        self.verbose = verbose
        self.locals = {}
        self.globals = {}
        self.parent = None

    def init_callback(self, agent):
        self.agent = agent
        self._init_callback()

    def _init_callback(self):
        pass

    def set_parent(self, parent):
        self.parent = parent

    def on_training_start(self, locals_, globals_):
        self.locals, self.globals = locals_, globals_
        self.num_time_steps = self.agent.num_time_steps
        self._on_training_start()

    def _on_training_start(self):
        pass

    def on_rollout_start(self):
        self._on_rollout_start()

    def _on_rollout_start(self):
        pass

    def _on_step(self):
        return True

    def on_step(self):
        self.n_calls += 1
        self.num_time_steps = self.agent.num_time_steps
        return self._on_step()

    def on_training_end(self):
        self._on_training_end()

    def _on_training_end(self):
        pass

    def on_rollout_end(self):
        """Returning False stops training (how callbacks that do not need per-step locals end a run: the device
        driver collects a whole rollout in one launch and never calls on_step for them)."""
        return self._on_rollout_end()

    def _on_rollout_end(self):
        return None

    def update_locals(self, locals_):
        self.locals.update(locals_)
        self.update_child_locals(locals_)

    def update_child_locals(self, locals_):
        pass

    # True when the callback never looks at per-step locals, so the whole rollout may run as one
    # device launch.
    needs_per_step = True


class CallbackList(BaseCallback):
    def __init__(self, callbacks):
        super().__init__()
        self.callbacks = list(callbacks)

    @property
    def needs_per_step(self):
        return any(getattr(c, "needs_per_step", True) for c in self.callbacks)

    def _init_callback(self):
        for c in self.callbacks:
            c.init_callback(self.agent)

    def _on_training_start(self):
        for c in self.callbacks:
            c.on_training_start(self.locals, self.globals)

    def _on_rollout_start(self):
        for c in self.callbacks:
            c.on_rollout_start()

    def _on_step(self):
        cont = True
        for c in self.callbacks:
            cont = c.on_step() and cont
        return cont

    def _on_rollout_end(self):
        cont = True
        for c in self.callbacks:
            cont = (c.on_rollout_end() is not False) and cont
        return cont

    def _on_training_end(self):
        for c in self.callbacks:
            c.on_training_end()

    def update_child_locals(self, locals_):
        for c in self.callbacks:
            c.update_locals(locals_)


class StopTrainingOnMaxSteps(BaseCallback):
    """Minimal stand-in used by tests: stops after max_calls on_step() calls."""

    def __init__(self, max_calls):
        super().__init__()
        self.max_calls = max_calls

    def _on_step(self):
        return self.n_calls < self.max_calls


class EvalCallback(BaseCallback):
    """Periodic evaluation on a second vector env (reference: openrl/utils/callbacks/eval_callback.py:53-284).

    `eval_freq` counts vector-env steps like the reference's `n_calls % eval_freq`; because this callback does not
    look at per-step locals the device driver keeps collecting whole rollouts in one launch and the evaluation runs
    at the first rollout boundary at or after each multiple of `eval_freq`.  Logs Eval/episode_reward(_std),
    Eval/episode_length(_std) through the agent's logger, appends to `<log_path>/evaluations.npz`, saves the best
    model to `<best_model_save_path>/best_model` and calls `callbacks_on_new_best` / `callbacks_after_eval`
    (their on_step() returning False stops training)."""

    needs_per_step = False

    def __init__(self, eval_env, callbacks_on_new_best=None, callbacks_after_eval=None, n_eval_episodes=5, eval_freq=10000,
                 log_path=None, best_model_save_path=None, deterministic=True, render=False, asynchronous=True, verbose=1,
                 warn=True, stop_logic="OR", close_env_at_end=True):
        super().__init__(verbose=verbose)
        if isinstance(eval_env, (str, dict)):
            from ..envs.common import make
            spec = {"id": eval_env, "env_num": 1} if isinstance(eval_env, str) else eval_env
            eval_env = make(spec["id"], env_num=spec.get("env_num", 1))
        self.eval_env = eval_env
        self.callbacks_on_new_best, self.callback = callbacks_on_new_best, callbacks_after_eval
        self.n_eval_episodes, self.eval_freq = n_eval_episodes, eval_freq
        self.deterministic, self.render, self.warn = deterministic, render, warn
        self.close_env_at_end = close_env_at_end
        self.best_mean_reward = self.last_mean_reward = -float("inf")
        self.best_model_save_path = best_model_save_path
        self.log_path = None if log_path is None else __import__("os").path.join(str(log_path), "evaluations")
        self.evaluations_results, self.evaluations_time_steps, self.evaluations_length = [], [], []
        self._evals_done = 0

    def _init_callback(self):
        import os
        if self.best_model_save_path is not None:
            os.makedirs(self.best_model_save_path, exist_ok=True)
        if self.log_path is not None:
            os.makedirs(os.path.dirname(self.log_path), exist_ok=True)
        for cb in (self.callbacks_on_new_best, self.callback):
            if cb is not None:
                cb.set_parent(self)
                cb.init_callback(self.agent)

    def _on_rollout_end(self):
        import os

        import numpy as np

        from .evaluation import evaluate_policy

        self.num_time_steps = self.agent.num_time_steps
        if self.eval_freq <= 0:
            return True
        vec_steps = self.num_time_steps // max(1, getattr(self.agent, "env_num", 1))
        due = vec_steps // self.eval_freq
        if due <= self._evals_done:
            return True
        self._evals_done = due
        rewards, lengths = evaluate_policy(self.agent, self.eval_env, n_eval_episodes=self.n_eval_episodes, render=self.render,
                                           deterministic=self.deterministic, return_episode_rewards=True, warn=self.warn)
        if self.log_path is not None:
            self.evaluations_time_steps.append(self.num_time_steps)
            self.evaluations_results.append(rewards)
            self.evaluations_length.append(lengths)
            np.savez(self.log_path, timesteps=self.evaluations_time_steps, results=self.evaluations_results,
                     ep_lengths=self.evaluations_length)
        mean_reward, std_reward = float(np.mean(rewards)), float(np.std(rewards))
        self.last_mean_reward = mean_reward
        info = {"Eval/episode_reward": mean_reward, "Eval/episode_reward_std": std_reward,
                "Eval/episode_length": float(np.mean(lengths)), "Eval/episode_length_std": float(np.std(lengths))}
        if self.verbose >= 1:
            print(f"Eval num_timesteps={self.num_time_steps}, episode_reward={mean_reward:.2f} +/- {std_reward:.2f}")
        cont = True
        if mean_reward > self.best_mean_reward:
            self.best_mean_reward = mean_reward
            if self.best_model_save_path is not None:
                self.agent.save(os.path.join(self.best_model_save_path, "best_model"))
                with open(os.path.join(self.best_model_save_path, "best_model_info.txt"), "w") as f:
                    f.write(f"best model at step: {self.num_time_steps}\n")
                    f.write(f"best model reward: {mean_reward}\n")
            if self.callbacks_on_new_best is not None:
                cont = self.callbacks_on_new_best.on_step() is not False
        if self.callback is not None:
            cont = (self.callback.on_step() is not False) and cont
        logger = getattr(self.agent, "logger", None)
        if logger is not None:
            logger.log_info(info, self.num_time_steps)
        return cont

    def _on_training_end(self):
        if self.close_env_at_end and hasattr(self.eval_env, "close"):
            self.eval_env.close()


class CheckpointCallback(BaseCallback):
    """agent.save(<save_path>/<name_prefix>_<num_time_steps>_steps) every `save_freq` vector-env steps
    (reference: utils/callbacks/checkpoint_callback.py:26-100).  Needs no per-step locals: the check runs at rollout
    boundaries, so the device driver keeps its one-launch rollouts."""

    needs_per_step = False

    def __init__(self, save_freq, save_path, name_prefix="rl_model", save_replay_buffer=False, verbose=0):
        super().__init__(verbose)
        if save_replay_buffer:
            raise NotImplementedError("on-policy agents have no replay buffer to checkpoint")
        self.save_freq, self.save_path, self.name_prefix = save_freq, str(save_path), name_prefix
        self._saves_done = 0

    def _init_callback(self):
        import os
        os.makedirs(self.save_path, exist_ok=True)

    def _on_rollout_end(self):
        import os
        self.num_time_steps = self.agent.num_time_steps
        due = (self.num_time_steps // max(1, getattr(self.agent, "env_num", 1))) // self.save_freq
        if due > self._saves_done:
            self._saves_done = due
            path = os.path.join(self.save_path, f"{self.name_prefix}_{self.num_time_steps}_steps")
            self.agent.save(path)
            if self.verbose >= 2:
                print(f"Saving model checkpoint to {path}")
        return True


class StopTrainingOnRewardThreshold(BaseCallback):
    """Child of EvalCallback (`callbacks_on_new_best`): stop once the best mean evaluation reward reaches the threshold
    (reference: utils/callbacks/stop_callback.py:23-53)."""

    def __init__(self, reward_threshold, verbose=0):
        super().__init__(verbose)
        self.reward_threshold = reward_threshold

    def _on_step(self):
        assert self.parent is not None, "StopTrainingOnRewardThreshold must be used with an EvalCallback"
        cont = bool(self.parent.best_mean_reward < self.reward_threshold)
        if self.verbose >= 1 and not cont:
            print(f"Stopping training because the mean reward {self.parent.best_mean_reward:.2f} is above the threshold "
                  f"{self.reward_threshold}")
        return cont


class StopTrainingOnNoModelImprovement(BaseCallback):
    """Child of EvalCallback (`callbacks_after_eval`): stop after more than `max_no_improvement_evals` consecutive
    evaluations without a new best mean reward, counting only after `min_evals` evaluations
    (reference: utils/callbacks/stop_callback.py:107-160)."""

    def __init__(self, max_no_improvement_evals, min_evals=0, verbose=0):
        super().__init__(verbose)
        self.max_no_improvement_evals, self.min_evals = max_no_improvement_evals, min_evals
        self.last_best_mean_reward = -float("inf")
        self.no_improvement_evals = 0

    def _on_step(self):
        assert self.parent is not None, "StopTrainingOnNoModelImprovement must be used with an EvalCallback"
        cont = True
        if self.n_calls > self.min_evals:
            if self.parent.best_mean_reward > self.last_best_mean_reward:
                self.no_improvement_evals = 0
            else:
                self.no_improvement_evals += 1
                cont = self.no_improvement_evals <= self.max_no_improvement_evals
        self.last_best_mean_reward = self.parent.best_mean_reward
        return cont


class StopTrainingOnMaxEpisodes(BaseCallback):
    """Stop after max_episodes * n_envs finished episodes, counted from the per-step `dones` local
    (reference: utils/callbacks/stop_callback.py:56-104) — so this one keeps the per-step launch mode."""

    needs_per_step = True

    def __init__(self, max_episodes, verbose=0):
        super().__init__(verbose)
        self.max_episodes, self.n_episodes, self._total = max_episodes, 0, max_episodes

    def _init_callback(self):
        self._total = self.max_episodes * getattr(self.agent, "env_num", 1)

    def _on_step(self):
        import numpy as np
        assert "dones" in self.locals, "`dones` is not among the step locals"
        d = np.asarray(self.locals["dones"])
        self.n_episodes += int(np.all(d.reshape(d.shape[0], -1), axis=1).sum())
        return self.n_episodes < self._total
