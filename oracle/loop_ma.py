"""Oracle: collect + update iteration for multi-agent envs with (optionally) recurrent policies —
MAPPO on MPE simple_spread as in examples/mpe (shared actor-critic over agents, GRU, chunked BPTT).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, with the reference's consumption order of
the global torch generator:
  OnPolicyDriver.act / add2buffer / compute_returns   openrl/drivers/onpolicy_driver.py:80-279
  ReplayData.insert / after_update                    openrl/buffers/replay_data.py:245-318
  ReplayData.recurrent_generator (chunks of L over the agent-major / time-minor flattening,
  ignoring trajectory boundaries)                     openrl/buffers/replay_data.py:1062-1258,
                                                      openrl/buffers/utils/util.py:88-97
  ReplayData.feed_forward_generator                   openrl/buffers/replay_data.py:553-646
  PPOAlgorithm.train_ppo / ppo_update                 openrl/algorithms/ppo.py:46-458
  RNNLayer                                            openrl/modules/networks/utils/rnn.py:39-99
Pinned against tests/golden/trace_mpe_gru.npz and trace_mpe_mlp.npz (tests/test_oracle_loop.py).
"""
import random
import types

import numpy as np
import torch

from . import envs as oenvs
from . import gae as ogae
from . import nets, ppo


def _obs_pair(obs):
    """Dict observation spaces feed policy / critic separately; plain spaces give both nets the same array."""
    return (obs["policy"], obs["critic"]) if isinstance(obs, dict) else (obs, obs)


def _cast(x):
    """(T, N, A, d) -> (N*A*T, d), agent-major / time-minor (buffers/utils/util.py:96-97)."""
    return x.transpose(1, 2, 0, 3).reshape(-1, *x.shape[3:])


class MATrainer:
    def __init__(self, cfg, env_id, env_num):
        self.cfg, self.N = cfg, env_num
        random.seed(cfg.seed)
        np.random.seed(cfg.seed)
        torch.manual_seed(cfg.seed)
        self.env = oenvs.ENVS[env_id](env_num)
        self.env.reset(seed=cfg.seed)
        A, d, n = self.env.agent_num, self.env.obs_dim, self.env.n_actions
        dc = getattr(self.env, "critic_obs_dim", d)
        self.A = A
        self.pol = nets.init_policy(cfg, d, "Discrete", n)
        self.cri = nets.init_critic(cfg, dc)
        self.opt_p, self.opt_c = ppo.make_optimizers(cfg, self.pol, self.cri)
        self.vn = ppo.ValueNormState() if cfg.use_valuenorm else None
        T, N, H = cfg.episode_length, env_num, cfg.hidden_size
        f = lambda *s: np.zeros(s, np.float32)  # noqa: E731
        one = lambda *s: np.ones(s, np.float32)  # noqa: E731
        self.buf = types.SimpleNamespace(
            policy_obs=f(T + 1, N, A, d), critic_obs=f(T + 1, N, A, dc), rnn_states=f(T + 1, N, A, 1, H),
            rnn_states_critic=f(T + 1, N, A, 1, H), value_preds=f(T + 1, N, A, 1), returns=f(T + 1, N, A, 1),
            masks=one(T + 1, N, A, 1), bad_masks=one(T + 1, N, A, 1), active_masks=one(T + 1, N, A, 1),
            action_masks=one(T + 1, N, A, n), actions=f(T, N, A, 1), action_log_probs=f(T, N, A, 1), rewards=f(T, N, A, 1))
        self.buf.policy_obs[0], self.buf.critic_obs[0] = _obs_pair(self.env.reset())

    def rollout(self):
        cfg, b, N, A = self.cfg, self.buf, self.N, self.A
        rec = nets.is_recurrent(cfg)
        for step in range(cfg.episode_length):
            with torch.no_grad():
                cat = lambda x: torch.from_numpy(np.concatenate(x))  # noqa: E731
                obs, cobs, masks, am = cat(b.policy_obs[step]), cat(b.critic_obs[step]), cat(b.masks[step]), cat(b.action_masks[step])
                hs, hc = (cat(b.rnn_states[step]), cat(b.rnn_states_critic[step])) if rec else (None, None)
                actions, logp, hs2 = nets.policy_act(self.pol, cfg, obs, am, hs, masks)
                values, hc2 = nets.critic_forward(self.cri, cfg, cobs, hc, masks)
            actions = actions.numpy().reshape(N, A, 1)
            obs2, rewards, dones, _ = self.env.step(actions)
            dones_env = np.all(dones, axis=1)
            if rec:
                hs2 = hs2.numpy().reshape(N, A, 1, -1).copy()
                hc2 = hc2.numpy().reshape(N, A, 1, -1).copy()
                hs2[dones_env] = 0.0
                hc2[dones_env] = 0.0
                b.rnn_states[step + 1] = hs2
                b.rnn_states_critic[step + 1] = hc2
            masks_n = np.ones((N, A, 1), np.float32)
            masks_n[dones_env] = 0.0
            active = np.ones((N, A, 1), np.float32)
            active[dones] = 0.0
            active[dones_env] = 1.0
            b.policy_obs[step + 1], b.critic_obs[step + 1] = _obs_pair(obs2)
            b.actions[step] = actions
            b.action_log_probs[step] = logp.numpy().reshape(N, A, 1)
            b.value_preds[step] = values.numpy().reshape(N, A, 1)
            b.rewards[step] = rewards
            b.masks[step + 1] = masks_n
            b.active_masks[step + 1] = active

    def compute_returns(self):
        cfg, b = self.cfg, self.buf
        with torch.no_grad():
            hc = torch.from_numpy(np.concatenate(b.rnn_states_critic[-1])) if nets.is_recurrent(cfg) else None
            nv, _ = nets.critic_forward(self.cri, cfg, torch.from_numpy(np.concatenate(b.critic_obs[-1])), hc,
                                        torch.from_numpy(np.concatenate(b.masks[-1])))
        nv = nv.numpy().reshape(self.N, self.A, 1)
        vn_state = self.vn.state() if self.vn is not None else None
        b.returns, b.value_preds = ogae.compute_returns(b.rewards, b.value_preds, b.masks, b.bad_masks, nv, cfg.gamma,
                                                        cfg.gae_lambda, cfg.use_gae, cfg.use_proper_time_limits, vn_state)

    def _recurrent_batches(self, adv):
        """recurrent_generator (replay_data.py:1062-1258)."""
        cfg, b = self.cfg, self.buf
        T, N, A = b.rewards.shape[:3]
        L = cfg.data_chunk_length
        batch_size = N * T * A
        data_chunks = batch_size // L
        mb = data_chunks // cfg.num_mini_batch
        rand = torch.randperm(data_chunks).numpy()
        flat = {k: _cast(getattr(b, k)[:T]) for k in ("policy_obs", "critic_obs", "actions", "action_log_probs", "value_preds",
                                                       "returns", "masks", "active_masks", "action_masks")}
        flat["adv"] = _cast(adv)
        hs = b.rnn_states[:-1].transpose(1, 2, 0, 3, 4).reshape(-1, *b.rnn_states.shape[3:])
        hc = b.rnn_states_critic[:-1].transpose(1, 2, 0, 3, 4).reshape(-1, *b.rnn_states_critic.shape[3:])
        for i in range(cfg.num_mini_batch):
            idx = rand[i * mb:(i + 1) * mb]
            out = {}
            for k, v in flat.items():
                st = np.stack([v[c * L:c * L + L] for c in idx], axis=1)  # (L, n, d)
                out[k] = torch.from_numpy(st.reshape(L * len(idx), *st.shape[2:]))
            out["rnn_states"] = torch.from_numpy(np.stack([hs[c * L] for c in idx]))
            out["rnn_states_critic"] = torch.from_numpy(np.stack([hc[c * L] for c in idx]))
            yield rand, out

    def _naive_batches(self, adv):
        """naive_recurrent_generator (replay_data.py:806-946): a minibatch is a set of (env, agent) rows with their WHOLE
        trajectories (T steps, time-major), initial hidden state = slot 0."""
        cfg, b = self.cfg, self.buf
        T, N, A = b.rewards.shape[:3]
        B = N * A
        per = B // cfg.num_mini_batch
        perm = torch.randperm(B).numpy()
        col = lambda x: x.reshape(x.shape[0], B, *x.shape[3:])   # noqa: E731
        arrs = {k: col(getattr(b, k)) for k in ("policy_obs", "critic_obs", "actions", "action_log_probs", "value_preds", "returns",
                                                 "masks", "active_masks", "action_masks", "rnn_states", "rnn_states_critic")}
        arrs["adv"] = col(adv)
        for start in range(0, B, per):
            ids = perm[start:start + per]
            if len(ids) < per:
                break
            out = {}
            for k in ("policy_obs", "critic_obs", "actions", "action_log_probs", "value_preds", "returns", "masks", "active_masks", "action_masks", "adv"):
                st = arrs[k][:T][:, ids]                       # (T, k, d): every step of the chosen rows
                out[k] = torch.from_numpy(np.ascontiguousarray(st).reshape(T * len(ids), *st.shape[2:]))
            out["rnn_states"] = torch.from_numpy(np.ascontiguousarray(arrs["rnn_states"][0][ids]))
            out["rnn_states_critic"] = torch.from_numpy(np.ascontiguousarray(arrs["rnn_states_critic"][0][ids]))
            yield perm, out

    def train(self):
        cfg, b = self.cfg, self.buf
        vn_state = self.vn.state() if self.vn is not None else None
        _, adv = ogae.advantages(b.returns, b.value_preds, b.active_masks, vn_state, cfg.use_adv_normalize)
        self.last_adv = adv
        updates, perms = [], []
        T, N, A = b.rewards.shape[:3]
        for _ in range(cfg.ppo_epoch):
            if cfg.use_recurrent_policy or getattr(cfg, "use_naive_recurrent_policy", False):
                gen = self._recurrent_batches(adv) if cfg.use_recurrent_policy else self._naive_batches(adv)   # ppo.py:363-381
                for rand, bt in gen:
                    batch = dict(critic_obs=bt["critic_obs"], policy_obs=bt["policy_obs"], actions=bt["actions"],
                                 value_preds=bt["value_preds"], returns=bt["returns"], active_masks=bt["active_masks"],
                                 old_logp=bt["action_log_probs"], adv=bt["adv"], action_masks=bt["action_masks"],
                                 masks=bt["masks"], rnn_states=bt["rnn_states"], rnn_states_critic=bt["rnn_states_critic"])
                    updates.append(ppo.ppo_update(cfg, self.pol, self.cri, self.opt_p, self.opt_c, self.vn, batch))
                perms.append(rand.copy())
            else:
                total = T * N * A
                mb = total // cfg.num_mini_batch
                rand = torch.randperm(total)
                perms.append(rand.numpy().copy())
                fl = lambda x: torch.from_numpy(np.ascontiguousarray(x).reshape(total, -1))  # noqa: E731
                for i in range(cfg.num_mini_batch):
                    idx = rand[i * mb:(i + 1) * mb]
                    batch = dict(critic_obs=fl(b.critic_obs[:-1])[idx], policy_obs=fl(b.policy_obs[:-1])[idx],
                                 actions=fl(b.actions)[idx], value_preds=fl(b.value_preds[:-1])[idx],
                                 returns=fl(b.returns[:-1])[idx], active_masks=fl(b.active_masks[:-1])[idx],
                                 old_logp=fl(b.action_log_probs)[idx], adv=fl(adv)[idx], action_masks=fl(b.action_masks[:-1])[idx])
                    updates.append(ppo.ppo_update(cfg, self.pol, self.cri, self.opt_p, self.opt_c, self.vn, batch))
        return np.array(updates, np.float64), np.stack(perms)

    def after_update(self):
        b = self.buf
        for name in ("policy_obs", "critic_obs", "rnn_states", "rnn_states_critic", "masks", "bad_masks", "active_masks", "action_masks"):
            getattr(b, name)[0] = getattr(b, name)[-1].copy()
