"""Device self-play path (BASELINE configs[3]): the two-player GridWorld step against the numpy oracle (bit-exact,
scripted actions for both players), opponent-pool sampling (RandomOpponent uniform over the ring / LastOpponent newest,
openrl/selfplay/sample_strategy/*.py), win/loss/draw bookkeeping, snapshot cadence, and a short training run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _env(N, table=None, **kw):
    from openrl_b200.envs.common import make

    return make("GridWorldSelfPlay", env_num=N, reset_table=table, **kw)


def _start_table(rng, N, K):
    t = np.zeros((N, K, 4), np.int64)
    for e in range(N):
        for k in range(K):
            while True:
                c = rng.integers(0, 10, 4)
                if tuple(c[:2]) != (1, 1) and tuple(c[2:]) != (1, 1) and tuple(c[:2]) != tuple(c[2:]):
                    t[e, k] = c
                    break
    return t


def test_two_player_gridworld_step_matches_oracle(cuda):
    import torch

    from openrl_b200 import lib
    from oracle.selfplay import GridWorld2P

    rng = np.random.default_rng(0)
    N, T, K = 64, 150, 40
    table = _start_table(rng, N, K)
    # biased scripted actions so that goals are reached often (moves towards (1,1) more likely) and time-outs occur too
    acts = rng.integers(0, 5, (T, N, 2))
    acts[:, : N // 4] = 0                                                  # a quarter of the envs never moves: 100-step time-outs
    env = _env(N, table)
    obs0, _ = env.reset(seed=0)
    ora = GridWorld2P(table)
    assert np.array_equal(obs0[:, 0, :], ora.reset())
    dev = env.device
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)   # noqa: E731
    buf = dict(obs=z(T + 1, N, 4), act=z(T, N), logp=z(T, N), rew=z(T, N), masks=torch.ones(T + 1, N, device=dev), active=torch.ones(T + 1, N, device=dev))
    buf["obs"][0].copy_(torch.from_numpy(obs0[:, 0, :]))
    scripted = torch.from_numpy(acts.astype(np.float32)).to(dev).contiguous()
    params = z(int(lib.load().orl_net_param_count(4, 5)))
    a = lib.OrlRolloutArgs()
    a.env_kind, a.n_envs, a.n_agents, a.episode_length = env.kind, N, 1, T
    a.t_begin, a.t_end, a.obs_dim, a.n_actions, a.activation_id, a.deterministic = 0, T, 4, 5, 1, 2 | 4
    a.policy_params, a.policy_obs = lib.ptr(params), lib.ptr(buf["obs"])
    a.actions, a.action_log_probs, a.rewards = lib.ptr(buf["act"]), lib.ptr(buf["logp"]), lib.ptr(buf["rew"])
    a.masks, a.active_masks, a.exp_noise = lib.ptr(buf["masks"]), lib.ptr(buf["active"]), lib.ptr(scripted)
    a.rng_seed = env.rng_seed
    a.env_i32, a.env_table, a.env_table_len = lib.ptr(env.env_i32), lib.ptr(env.env_table), env.env_table_len
    a.ep_return, a.ep_length, a.episode_stats = lib.ptr(env.ep_return), lib.ptr(env.ep_length), lib.ptr(env.episode_stats)
    lib.check(lib.load().orl_selfplay_rollout(env.selfplay_args(a), lib.current_stream()), "rollout")
    torch.cuda.synchronize()
    obs, rew, masks = buf["obs"].cpu().numpy(), buf["rew"].cpu().numpy(), buf["masks"].cpu().numpy()
    for t in range(T):
        o, r, d = ora.step(acts[t, :, 0], acts[t, :, 1])
        assert np.array_equal(obs[t + 1], o), t
        assert np.array_equal(rew[t], r), t
        assert np.array_equal(masks[t + 1] == 0, d), t
    assert np.array_equal(buf["act"].cpu().numpy(), acts[:, :, 0].astype(np.float32))
    # bookkeeping: wins / losses / draws against the random-action opponent slot (the pool is empty) == the oracle's tally
    st = env.opponent_pool.stats.cpu().numpy()
    assert np.array_equal(st[-1], ora.outcomes) and st[:-1].sum() == 0 and ora.outcomes.min() > 0
    assert int(env.episode_stats.cpu().numpy()[2]) == ora.outcomes.sum()


@pytest.mark.parametrize("strategy", ["RandomOpponent", "LastOpponent"])
def test_opponent_sampling_strategies(cuda, strategy):
    import torch
    from scipy import stats

    N, cap = 4096, 4
    env = _env(N, opponent_pool_size=cap, opponent_strategy=strategy)
    pool = env.opponent_pool
    flat = torch.zeros(pool.stride, device=env.device)
    for k in range(6):                     # 6 snapshots into a ring of 4: slots hold snapshots 4, 5, 2, 3
        pool.add(flat + k)
    assert pool.count == 6 and int(pool.count_dev.item()) == 6
    assert [float(pool.params[s, 0]) for s in range(cap)] == [4.0, 5.0, 2.0, 3.0]
    env.reset(seed=5)
    opp = env.env_i32[6].cpu().numpy()
    if strategy == "LastOpponent":
        assert (opp == (6 - 1) % cap).all()                                  # the newest snapshot (last_opponent.py:24-27)
    else:
        counts = np.bincount(opp, minlength=cap)
        assert counts.min() > 0 and stats.chisquare(counts).pvalue > 1e-4      # uniform over the ring (random_opponent.py:25-28)
    # empty pool: random-action opponent
    env2 = _env(16, opponent_pool_size=cap, opponent_strategy=strategy)
    env2.reset(seed=1)
    assert (env2.env_i32[6].cpu().numpy() == -1).all()


def test_selfplay_training_runs_and_snapshots(cuda):
    from openrl_b200.configs.config import create_config_parser
    from openrl_b200.modules.common import PPONet
    from openrl_b200.runners.common import PPOAgent
    from openrl_b200.utils.logger import Logger

    T, N, iters = 64, 256, 9
    cfg = create_config_parser().parse_args(["--seed", "0", "--episode_length", str(T), "--ppo_epoch", "2", "--log_interval", "1",
                                             "--selfplay_save_freq", "2"])
    cfg.quiet = True
    env = _env(N, opponent_pool_size=3)
    agent = PPOAgent(PPONet(env, cfg=cfg, device="cuda:0"))
    logger = Logger(quiet=True)
    agent.train(total_time_steps=T * N * iters, logger=logger)
    pool = env.opponent_pool
    assert pool.count == iters // 2                                   # snapshots after iterations 2, 4, 6, 8
    assert getattr(agent.driver, "_graph", None) is not None           # the self-play iteration is graph-replayed too
    logs = [h[1] for h in logger.history if "value_loss" in h[1]]
    assert len(logs) == iters and all(np.isfinite(list(l.values())).all() for l in logs)
    res = pool.battle_results()
    played = sum(sum(v) for v in res.values())
    assert played > 0 and res["random"][0] + res["random"][1] + res["random"][2] > 0
    # the newest snapshot equals the policy as it was after iteration 8 ... and differs from the current one only by iteration 9
    import torch

    pol = agent.net.module.models["policy"].flat_params
    newest = pool.params[(pool.count - 1) % pool.capacity, :pol.numel()]
    assert not torch.equal(newest, pol) and float((newest - pol).abs().max()) < 0.05
    # the vec-env step API works on the same env (learner actions given, opponent from the pool)
    obs, _ = env.reset(seed=3)
    o2, r, d, info = env.step(np.zeros((N, 1, 1)))
    assert o2.shape == (N, 1, 4) and r.shape == (N, 1, 1) and d.shape == (N, 1)
    keep = ~d[:, 0]                                                    # envs whose episode did not end in this step
    assert keep.any() and (o2[keep, 0, :2] == obs[keep, 0, :2]).all()  # action 0 = stay: the learner did not move
