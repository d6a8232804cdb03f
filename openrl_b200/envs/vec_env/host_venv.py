"""Host-stepped vec-env adapter (BASELINE config 5): the env.step stays on the
host (MuJoCo-class simulators, the reference's own SyncVectorEnv/AsyncVectorEnv, ...) while the
policy forward, buffer, GAE and update run on the device.

Wraps any object with the reference's BaseVecEnv duck type — `reset(seed=) -> obs | (obs, infos)`,
`step(actions) -> (obs (N,A,d), rewards (N,A,1), dones (N,A), infos)`, `parallel_env_num`,
`agent_num`, `observation_space`, `action_space` — and adds pinned staging buffers so that each step
costs one D2H copy (actions) and one H2D copy (obs, rewards, dones).  `kind = ORL_ENV_NONE` tells
the driver to run the per-step loop (onpolicy_driver.py:154-203 semantics)."""
import numpy as np
import torch

from ... import lib


class HostVecEnv:
    def __init__(self, env, device="cuda:0"):
        self.env = env
        self.kind = lib.ENV_NONE
        self.device = torch.device(device)
        self.parallel_env_num = env.parallel_env_num
        self.agent_num = env.agent_num
        self.observation_space = env.observation_space
        self.action_space = env.action_space
        self.env_name = getattr(env, "env_name", type(env).__name__)
        self.use_monitor = False
        self.env_table, self.env_table_len = None, 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        N, A = self.parallel_env_num, self.agent_num
        if not hasattr(self.observation_space, "shape") or self.observation_space.shape is None:
            raise NotImplementedError("HostVecEnv stages flat Box observations; Dict observation spaces are not supported")
        d = self.observation_space.shape[0]
        self.obs_dim = d
        self._discrete = hasattr(self.action_space, "n")
        self._stage = torch.empty(N * A * (d + 2), dtype=torch.float32, pin_memory=torch.cuda.is_available())

    def reset(self, seed=None, options=None):
        out = self.env.reset(seed=seed) if seed is not None else self.env.reset()
        if isinstance(out, tuple):
            return out
        return out, [{} for _ in range(self.parallel_env_num)]

    def reset_into(self, obs_out, critic_obs_out=None):
        obs, _ = self.reset()
        obs_out.copy_(torch.as_tensor(np.asarray(obs, dtype=np.float32)).view_as(obs_out), non_blocking=False)

    def step(self, actions, extra_data=None):
        return self.env.step(actions)

    def step_staged(self, actions_dev):
        """actions (B, w) device tensor -> host env.step -> the device block [obs | rewards | dones] (one pinned H2D copy,
        the layout orl_host_insert reads) plus the host-side step outputs."""
        N, A, d = self.parallel_env_num, self.agent_num, self.obs_dim
        a = actions_dev.cpu().numpy().reshape(N, A, -1)          # D2H (synchronises the stream)
        self.d2h_bytes += a.nbytes
        if self._discrete:                                       # the reference hands integer indices to env.step
            a = a.astype(np.int64)
        obs, rewards, dones, infos = self.env.step(a)
        st = self._stage.numpy()
        B = N * A
        st[:B * d] = np.asarray(obs, dtype=np.float32).reshape(-1)
        st[B * d:B * d + B] = np.asarray(rewards, dtype=np.float32).reshape(-1)
        st[B * d + B:] = np.asarray(dones, dtype=np.float32).reshape(-1)
        dev = self._stage.to(self.device, non_blocking=True)      # one H2D copy
        self.h2d_bytes += st.nbytes
        return dev, obs, rewards, dones, infos

    # -- two-group ping-pong (double-buffered ingest) ---------------------------------------------
    @property
    def supports_groups(self):
        """True when the wrapped vec-env can step a sub-range of its envs (`step_range(lo, hi, actions)`)."""
        return hasattr(self.env, "step_range") and self.parallel_env_num >= 2

    def group_bounds(self, n_groups=2):
        N = self.parallel_env_num
        cuts = [N * g // n_groups for g in range(n_groups + 1)]
        return [(cuts[g], cuts[g + 1]) for g in range(n_groups)]

    def _group_stage(self, g, lo, hi):
        key = (g, lo, hi)
        if getattr(self, "_gstages", None) is None:
            self._gstages = {}
        if key not in self._gstages:
            n, A, d = hi - lo, self.agent_num, self.obs_dim
            pin = torch.cuda.is_available()
            self._gstages[key] = dict(
                inp=torch.empty(n * A * (d + 2), dtype=torch.float32, pin_memory=pin),     # obs | rewards | dones
                act=None, dev=torch.empty(n * A * (d + 2), dtype=torch.float32, device=self.device),
                ev=torch.cuda.Event() if pin else None)
        return self._gstages[key]

    def group_fetch_actions(self, g, lo, hi, actions_dev):
        """Enqueue the D2H copy of this group's actions into its pinned buffer and mark it with an event."""
        st = self._group_stage(g, lo, hi)
        if st["act"] is None or st["act"].shape != actions_dev.shape:
            st["act"] = torch.empty(actions_dev.shape, dtype=torch.float32, pin_memory=torch.cuda.is_available())
        st["act"].copy_(actions_dev, non_blocking=True)
        if st["ev"] is not None:
            st["ev"].record()
        self.d2h_bytes += actions_dev.numel() * 4

    def group_step(self, g, lo, hi):
        """Wait for the group's actions, step its envs on the host, stage the results in pinned memory and enqueue ONE
        H2D copy; returns the device block [obs | rewards | dones] (the layout orl_host_insert reads) plus the host-side
        step outputs."""
        st = self._group_stage(g, lo, hi)
        if st["ev"] is not None:
            st["ev"].synchronize()
        n, A, d = hi - lo, self.agent_num, self.obs_dim
        a = st["act"].numpy().reshape(n, A, -1)
        if self._discrete:
            a = a.astype(np.int64)
        obs, rewards, dones, infos = self.env.step_range(lo, hi, a)
        B = n * A
        buf = st["inp"].numpy()
        buf[:B * d] = np.asarray(obs, dtype=np.float32).reshape(-1)
        buf[B * d:B * d + B] = np.asarray(rewards, dtype=np.float32).reshape(-1)
        buf[B * d + B:] = np.asarray(dones, dtype=np.float32).reshape(-1)
        st["dev"].copy_(st["inp"], non_blocking=True)
        self.h2d_bytes += buf.nbytes
        dev = st["dev"]
        return dev, obs, rewards, dones, infos

    def random_action(self, infos=None):
        return np.array([[self.action_space.sample() for _ in range(self.agent_num)] for _ in range(self.parallel_env_num)])

    # call / exec_func / set_attr (base_venv.py:231-302): forwarded to the wrapped host vec-env
    def call(self, name, *args, **kwargs):
        return self.env.call(name, *args, **kwargs)

    def get_attr(self, name):
        return self.env.call(name)

    def set_attr(self, name, values):
        return self.env.set_attr(name, values)

    def exec_func(self, func, indices=None, *args, **kwargs):
        return self.env.exec_func(func, indices, *args, **kwargs)

    def batch_rewards(self, buffer):
        return {}

    def statistics(self, buffer):
        return {}

    def close(self):
        if hasattr(self.env, "close"):
            self.env.close()
