"""Oracle env restatements against the reference traces: replaying the reference's recorded actions
must reproduce its observations / rewards / masks."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import envs as oenvs


def test_simple_spread_replay_matches_reference():
    d = np.load(os.path.join(GOLDEN, "trace_mpe_gru.npz"), allow_pickle=True)
    N = int(d["meta/env_num"])
    env = oenvs.SimpleSpreadVec(N)
    env.reset(seed=0)          # PPONet.__init__
    obs = env.reset()          # RLDriver.reset_and_buffer_init
    for it in range(int(d["meta/iters"])):
        pol, cri = d[f"it{it}/policy_obs"], d[f"it{it}/critic_obs"]
        acts, rews, masks = d[f"it{it}/actions"], d[f"it{it}/rewards"], d[f"it{it}/masks"]
        np.testing.assert_array_equal(obs["policy"].astype(np.float32), pol[0])
        np.testing.assert_array_equal(obs["critic"].astype(np.float32), cri[0])
        for t in range(acts.shape[0]):
            obs, r, done, _ = env.step(acts[t])
            np.testing.assert_array_equal(obs["policy"].astype(np.float32), pol[t + 1])
            np.testing.assert_array_equal(obs["critic"].astype(np.float32), cri[t + 1])
            np.testing.assert_array_equal(r.astype(np.float32), rews[t])
            np.testing.assert_array_equal((~done[:, :, None]).astype(np.float32), masks[t + 1])


def test_gridworld_replay_matches_reference():
    d = np.load(os.path.join(GOLDEN, "trace_gridworld.npz"), allow_pickle=True)
    N = int(d["meta/env_num"])
    # replay the reference's reset cells in global order (it drew them from the global MT19937)
    cells = []
    obs0 = d["it0/policy_obs"][0]
    cells += [obs0[e, 0, :2].astype(np.int64) for e in range(N)]
    for it in range(int(d["meta/iters"])):
        obs, masks = d[f"it{it}/policy_obs"], d[f"it{it}/masks"]
        for t in range(1, obs.shape[0]):
            for e in range(N):
                if masks[t, e, 0, 0] == 0.0:
                    cells.append(obs[t, e, 0, :2].astype(np.int64))
    env = oenvs.GridWorldVec(N, reset_table=np.array(cells))
    o = env.reset()
    for it in range(int(d["meta/iters"])):
        pol, acts, rews = d[f"it{it}/policy_obs"], d[f"it{it}/actions"], d[f"it{it}/rewards"]
        np.testing.assert_array_equal(o.astype(np.float32), pol[0])
        for t in range(acts.shape[0]):
            o, r, done, _ = env.step(acts[t].astype(np.int64))
            np.testing.assert_array_equal(o.astype(np.float32), pol[t + 1])
            np.testing.assert_array_equal(r.astype(np.float32), rews[t])
