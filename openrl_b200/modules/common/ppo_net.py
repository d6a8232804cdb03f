"""PPONet(env, cfg=None, device=..., n_rollout_threads=1, model_dict=None, module_class=PPOModule)
(reference: openrl/modules/common/ppo_net.py:50-143)."""
import numpy as np
import torch

from ...configs.config import create_config_parser
from ...utils.util import set_seed
from ..ppo_module import PPOModule


class PPONet:
    def __init__(self, env, cfg=None, device="cuda:0", n_rollout_threads=1, model_dict=None, module_class=PPOModule):
        if cfg is None:
            cfg = create_config_parser().parse_args([])
        set_seed(cfg.seed)
        env.reset(seed=cfg.seed)
        cfg.num_agents = env.agent_num
        cfg.n_rollout_threads = n_rollout_threads
        cfg.learner_n_rollout_threads = cfg.n_rollout_threads
        if cfg.rnn_type not in ("gru", "lstm"):
            raise NotImplementedError(f"RNN type {cfg.rnn_type} has not been implemented.")
        cfg.rnn_hidden_size = cfg.hidden_size if cfg.rnn_type == "gru" else cfg.hidden_size * 2
        if isinstance(device, str):
            device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("openrl_b200 runs on CUDA only (there is no CPU path)")
        self.module = module_class(cfg=cfg, policy_input_space=env.observation_space,
                                   critic_input_space=env.observation_space, act_space=env.action_space,
                                   share_model=cfg.use_share_model, device=device, rank=0, world_size=1,
                                   model_dict=model_dict)
        self.cfg, self.env, self.device = cfg, env, device
        self.rnn_states_actor, self.masks = None, None

    def act(self, observation, action_masks=None, deterministic=False, episode_starts=None):
        if self.cfg.use_recurrent_policy or self.cfg.use_naive_recurrent_policy:
            if episode_starts is not None and self.rnn_states_actor is not None:
                # reset_rnn_states (ppo_net.py:33-47): zero the hidden state of every agent of a restarted env
                keep = 1.0 - np.repeat(np.asarray(episode_starts, dtype=np.float32), self.env.agent_num)
                keep = torch.as_tensor(keep, dtype=torch.float32).to(self.device)[:, None, None]
                self.rnn_states_actor = torch.as_tensor(self.rnn_states_actor, dtype=torch.float32).to(self.device) * keep
            actions, _, self.rnn_states_actor = self.module.act(
                obs=observation, rnn_states_actor=self.rnn_states_actor, masks=self.masks, action_masks=action_masks,
                deterministic=deterministic)
            return actions, self.rnn_states_actor
        actions, _ = self.module.act(obs=observation, rnn_states_actor=self.rnn_states_actor, masks=self.masks,
                                     action_masks=action_masks, deterministic=deterministic)
        return actions, self.rnn_states_actor

    def reset(self, env=None):
        if env is not None:
            self.env = env
        self.rnn_states_actor, self.masks = self.module.init_rnn_states(
            rollout_num=self.env.parallel_env_num, agent_num=self.env.agent_num, rnn_layers=self.cfg.recurrent_N,
            hidden_size=self.cfg.rnn_hidden_size)

    def load_policy(self, path):
        self.module.load_policy(path)
