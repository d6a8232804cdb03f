"""Oracle: the whole collect + update iteration (feed-forward MLP, Discrete actions).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, with the same consumption order of
the global torch generator as the reference (param init -> one exponential_ per rollout step
-> one randperm per epoch):
  PPONet.__init__                    openrl/modules/common/ppo_net.py:50-96
  RLDriver.run / reset_and_buffer_init   openrl/drivers/rl_driver.py:118-157
  OnPolicyDriver.actor_rollout/add2buffer/act/compute_returns
                                     openrl/drivers/onpolicy_driver.py:80-279
  ReplayData.insert/after_update/feed_forward_generator
                                     openrl/buffers/replay_data.py:245-318,553-646
  PPOAlgorithm.train_ppo             openrl/algorithms/ppo.py:383-458
Pinned against tests/golden/trace_cartpole*.npz / trace_gridworld.npz (tests/test_oracle_loop.py).
"""
import random
import types

import numpy as np
import torch

from . import envs as oenvs
from . import gae as ogae
from . import nets, ppo

DEFAULTS = dict(
    seed=0, episode_length=200, hidden_size=64, layer_N=1, activation_id=1, use_feature_normalization=False,
    use_orthogonal=True, gain=0.01, use_recurrent_policy=False, recurrent_N=1, data_chunk_length=2,
    lr=5e-4, critic_lr=5e-4, opti_eps=1e-5, weight_decay=0.0, ppo_epoch=10, use_clipped_value_loss=True,
    clip_param=0.2, num_mini_batch=1, entropy_coef=0.01, value_loss_coef=0.5, use_max_grad_norm=True,
    max_grad_norm=10.0, use_gae=True, gamma=0.99, gae_lambda=0.95, use_proper_time_limits=False,
    use_huber_loss=True, huber_delta=10.0, use_value_active_masks=True, use_policy_active_masks=True,
    use_adv_normalize=False, use_valuenorm=True, use_popart=False, dual_clip_ppo=False, dual_clip_coeff=3.0,
    a2c=False, use_share_model=False, use_naive_recurrent_policy=False, use_linear_lr_decay=False,
)


def make_cfg(**kw):
    d = dict(DEFAULTS)
    d.update(kw)
    return types.SimpleNamespace(**d)


def cfg_from_flags(flag_string):
    toks = flag_string.split()
    kw = {}
    for k, v in zip(toks[::2], toks[1::2]):
        k = k.lstrip("-")
        if k not in DEFAULTS:
            continue
        t = type(DEFAULTS[k])
        kw[k] = (v.lower() in ("true", "1")) if t is bool else t(v)
    return make_cfg(**kw)


class Trainer:
    def __init__(self, cfg, env_id, env_num, env=None):
        self.cfg, self.N = cfg, env_num
        # PPONet.__init__: set_seed, env.reset(seed), build policy then critic
        random.seed(cfg.seed)
        np.random.seed(cfg.seed)
        torch.manual_seed(cfg.seed)
        self.env = env if env is not None else oenvs.ENVS[env_id](env_num)
        self.env.reset(seed=cfg.seed)
        d = self.env.obs_dim
        self.box = hasattr(self.env, "act_dim")
        n = self.env.act_dim if self.box else self.env.n_actions
        if getattr(cfg, "use_share_model", False):
            self.pol = self.cri = nets.init_policy_value(cfg, d, "Box" if self.box else "Discrete", n)
        else:
            self.pol = nets.init_policy(cfg, d, "Box" if self.box else "Discrete", n)
            self.cri = nets.init_critic(cfg, d)
        self.opt_p, self.opt_c = ppo.make_optimizers(cfg, self.pol, self.cri)
        self.vn = ppo.ValueNormState() if cfg.use_valuenorm else None   # base_value_network.py:31-34 (use_popart adds no normaliser)
        T, N, A = cfg.episode_length, env_num, 1
        f = lambda *s: np.zeros(s, np.float32)
        self.buf = types.SimpleNamespace(
            obs=f(T + 1, N, A, d), value_preds=f(T + 1, N, A, 1), returns=f(T + 1, N, A, 1),
            masks=np.ones((T + 1, N, A, 1), np.float32), bad_masks=np.ones((T + 1, N, A, 1), np.float32),
            active_masks=np.ones((T + 1, N, A, 1), np.float32), action_masks=np.ones((T + 1, N, A, n), np.float32),
            actions=f(T, N, A, n if self.box else 1), action_log_probs=f(T, N, A, n if self.box else 1),
            rewards=f(T, N, A, 1))
        # RLDriver.reset_and_buffer_init: a second, unseeded reset
        self.buf.obs[0] = self.env.reset()
        self.log = []

    def rollout(self):
        cfg, b = self.cfg, self.buf
        for step in range(cfg.episode_length):
            with torch.no_grad():
                obs = torch.from_numpy(np.concatenate(b.obs[step]))
                am = torch.from_numpy(np.concatenate(b.action_masks[step]))
                if self.box:
                    actions, logp = nets.policy_act_gaussian(self.pol, cfg, obs)
                else:
                    actions, logp, _ = nets.policy_act(self.pol, cfg, obs, am)
                values, _ = nets.critic_forward(self.cri, cfg, obs)
            actions = actions.numpy().reshape(self.N, 1, -1)
            obs2, rewards, dones, _ = self.env.step(actions)
            dones_env = np.all(dones, axis=1)
            masks = np.ones((self.N, 1, 1), np.float32)
            masks[dones_env] = 0.0
            active = np.ones((self.N, 1, 1), np.float32)
            active[dones] = 0.0
            active[dones_env] = 1.0
            b.obs[step + 1] = obs2
            b.actions[step] = actions
            b.action_log_probs[step] = logp.numpy().reshape(self.N, 1, -1)
            b.value_preds[step] = values.numpy().reshape(self.N, 1, 1)
            b.rewards[step] = rewards
            b.masks[step + 1] = masks
            b.active_masks[step + 1] = active

    def lr_decay(self, episode, episodes):
        """RLDriver.learner_update (rl_driver.py:159-161) -> PPOModule.lr_decay (ppo_module.py:91-100) ->
        update_linear_schedule (modules/utils/util.py:13-17): lr = lr0 - lr0 * episode / episodes, before the returns."""
        cfg = self.cfg
        opts = [(self.opt_p, cfg.lr)] if self.opt_c is self.opt_p else [(self.opt_p, cfg.lr), (self.opt_c, cfg.critic_lr)]
        for opt, lr0 in opts:
            lr = lr0 - (lr0 * (episode / float(episodes)))
            for group in opt.param_groups:
                group["lr"] = lr

    def compute_returns(self):
        cfg, b = self.cfg, self.buf
        with torch.no_grad():
            nv, _ = nets.critic_forward(self.cri, cfg, torch.from_numpy(np.concatenate(b.obs[-1])))
        nv = nv.numpy().reshape(self.N, 1, 1)
        vn_state = self.vn.state() if self.vn is not None else None
        b.returns, b.value_preds = ogae.compute_returns(
            b.rewards, b.value_preds, b.masks, b.bad_masks, nv, cfg.gamma, cfg.gae_lambda, cfg.use_gae,
            cfg.use_proper_time_limits, vn_state)

    def train(self):
        cfg, b = self.cfg, self.buf
        vn_state = self.vn.state() if self.vn is not None else None
        _, adv = ogae.advantages(b.returns, b.value_preds, b.active_masks, vn_state, cfg.use_adv_normalize)
        self.last_adv = adv
        T, N, A = b.rewards.shape[:3]
        batch_size = T * N * A
        mb = batch_size // cfg.num_mini_batch
        flat = lambda x: torch.from_numpy(x.reshape(batch_size, -1))
        obs, actions = flat(b.obs[:-1]), flat(b.actions)
        vp, ret = flat(b.value_preds[:-1]), flat(b.returns[:-1])
        active, logp, advf = flat(b.active_masks[:-1]), flat(b.action_log_probs), flat(adv)
        am = flat(b.action_masks[:-1])
        updates, perms = [], []
        for _ in range(cfg.ppo_epoch):
            rand = torch.randperm(batch_size)
            perms.append(rand.numpy().copy())
            for i in range(cfg.num_mini_batch):
                idx = rand[i * mb:(i + 1) * mb]
                batch = dict(critic_obs=obs[idx], policy_obs=obs[idx], actions=actions[idx], value_preds=vp[idx],
                             returns=ret[idx], active_masks=active[idx], old_logp=logp[idx], adv=advf[idx],
                             action_masks=am[idx])
                updates.append(ppo.ppo_update(cfg, self.pol, self.cri, self.opt_p, self.opt_c, self.vn, batch))
        return np.array(updates, np.float64), np.stack(perms)

    def after_update(self):
        b = self.buf
        for name in ("obs", "masks", "bad_masks", "active_masks", "action_masks"):
            getattr(b, name)[0] = getattr(b, name)[-1].copy()

    def iteration(self, episode=0, episodes=1):
        self.rollout()
        if self.cfg.use_linear_lr_decay:
            self.lr_decay(episode, episodes)
        self.compute_returns()
        out = self.train()
        self.after_update()
        return out
