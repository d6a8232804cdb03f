"""Statistical checks of the device (Philox4x32-10) samplers used outside parity mode: the sampled actions
must follow the policy's distribution (reference: Categorical.sample -> torch.multinomial, act.py:79-81,
distributions.py:16-33; DiagGaussian Normal.sample, distributions.py:34-47), successive calls must draw
fresh noise, and the fused rollout's draws must be independent across steps."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _net(act_space, obs_dim=4, seed=1):
    from openrl_b200 import spaces
    from openrl_b200.configs.config import create_config_parser
    from openrl_b200.modules.common import PPONet

    class Env:
        agent_num, parallel_env_num = 1, 1
        observation_space, action_space = spaces.Box(-5, 5, (obs_dim,), np.float32), act_space

        def reset(self, seed=None):
            return np.zeros((1, 1, obs_dim), np.float32)

    cfg = create_config_parser().parse_args(["--seed", str(seed)])
    cfg.quiet = True
    return cfg, PPONet(Env(), cfg=cfg, device="cuda:0")


def _oracle_params(net):
    import torch

    return {k: torch.from_numpy(v.detach().cpu().numpy().copy()) for k, v in net.module.models["policy"].state_dict().items()}


def _chi2_sf(x, k):
    from scipy import stats

    return float(stats.chi2.sf(x, k))


@pytest.mark.parametrize("n_actions", [2, 5])
def test_categorical_philox_sampling_follows_policy_probs(cuda, n_actions):
    import torch

    from openrl_b200 import spaces
    from oracle import nets

    cfg, net = _net(spaces.Discrete(n_actions))
    # sharpen the head a little so the probabilities are far from uniform
    sd = net.module.models["policy"].state_dict()
    sd["act.action_out.linear.weight"].mul_(40.0)
    p = _oracle_params(net)
    rows = 400_000
    pvals = []
    for obs_vec in ([0.3, -1.2, 0.7, 2.0], [-2.0, 0.1, 0.0, 1.0]):
        obs = np.tile(np.asarray(obs_vec, np.float32), (rows, 1))
        feat, _ = nets.policy_features(p, cfg, torch.from_numpy(obs[:1]))
        logits = nets.categorical_logits(p, feat)
        probs = torch.softmax(logits, -1)[0].numpy().astype(np.float64)
        assert probs.min() > 1e-3 and probs.max() < 0.95, probs
        counts = np.zeros(n_actions)
        for call in range(2):   # two calls: the call counter advances the Philox step
            actions, logp = net.module.act(obs, deterministic=False)
            a = actions.cpu().numpy().astype(np.int64).ravel()
            counts += np.bincount(a, minlength=n_actions)
            # the returned log-prob is the log-probability of the sampled action
            np.testing.assert_allclose(np.exp(logp.cpu().numpy().ravel()[:1000]), probs[a[:1000]], rtol=2e-4)
            if call == 0:
                first = a.copy()
        assert (first != a).mean() > 0.05, "a second act() call must see fresh noise"
        expected = probs * counts.sum()
        chi2 = float(((counts - expected) ** 2 / expected).sum())
        pvals.append(_chi2_sf(chi2, n_actions - 1))
    assert min(pvals) > 1e-4, pvals     # 800k draws per observation: a biased sampler gives p ~ 0


def test_rollout_kernel_draws_are_independent_across_steps_and_rows(cuda):
    """Fused CartPole rollout with device noise: with a policy whose logits are ~0 the action marginal is
    ~1/2 per (step, env); lag-1 autocorrelation over steps and cross-env correlation must vanish."""
    import torch

    from openrl_b200.configs.config import create_config_parser
    from openrl_b200.envs.common import make
    from openrl_b200.modules.common import PPONet
    from openrl_b200.runners.common import PPOAgent

    T, N = 64, 4096
    cfg = create_config_parser().parse_args(["--seed", "3", "--episode_length", str(T), "--ppo_epoch", "1", "--lr", "0", "--critic_lr", "0"])
    cfg.quiet = True
    env = make("CartPole-v1", env_num=N)
    net = PPONet(env, cfg=cfg, device="cuda:0")
    net.module.models["policy"].state_dict()["act.action_out.linear.weight"].zero_()
    agent = PPOAgent(net)
    agent.train(total_time_steps=T * N)
    a = agent.driver.buffer.data.actions.cpu().numpy().reshape(T, N)
    assert abs(a.mean() - 0.5) < 4 * 0.5 / np.sqrt(T * N)
    z = a - a.mean()
    lag1 = float((z[1:] * z[:-1]).mean() / z.var())
    cross = float((z[:, 1:] * z[:, :-1]).mean() / z.var())
    assert abs(lag1) < 5 / np.sqrt(T * N) and abs(cross) < 5 / np.sqrt(T * N), (lag1, cross)
    lp = agent.driver.buffer.data.action_log_probs.cpu().numpy()
    np.testing.assert_allclose(lp, np.log(0.5), atol=1e-5)


def test_gaussian_philox_sampling_moments(cuda):
    import torch

    from openrl_b200 import spaces
    from oracle import nets

    cfg, net = _net(spaces.Box(-1, 1, (3,), np.float32))
    sd = net.module.models["policy"].state_dict()
    sd["act.action_out.logstd._bias"].copy_(torch.tensor([[-0.5], [0.0], [0.4]]))
    p = _oracle_params(net)
    rows = 400_000
    obs = np.tile(np.asarray([0.5, -0.5, 1.5, 0.2], np.float32), (rows, 1))
    feat, _ = nets.policy_features(p, cfg, torch.from_numpy(obs[:1]))
    mean, std = nets.gaussian_params(p, feat)
    mean, std = mean[0].numpy().astype(np.float64), std.reshape(-1).numpy().astype(np.float64)
    actions, logp = net.module.act(obs, deterministic=False)
    a = actions.cpu().numpy().astype(np.float64)
    z = (a - mean) / std
    se = 1 / np.sqrt(rows)
    assert np.all(np.abs(z.mean(0)) < 5 * se), z.mean(0)
    assert np.all(np.abs(z.var(0) - 1) < 5 * np.sqrt(2) * se), z.var(0)
    assert np.all(np.abs((z ** 3).mean(0)) < 5 * np.sqrt(15) * se)            # skewness
    assert np.all(np.abs((z ** 4).mean(0) - 3) < 5 * np.sqrt(96) * se)        # kurtosis
    c = np.corrcoef(z.T)
    assert np.all(np.abs(c - np.eye(3)) < 5 * se)                             # dimensions are independent
    want_lp = -0.5 * z[:1000] ** 2 - np.log(std) - 0.5 * np.log(2 * np.pi)
    np.testing.assert_allclose(logp.cpu().numpy()[:1000], want_lp, rtol=2e-4, atol=2e-5)
    a2, _ = net.module.act(obs, deterministic=False)
    assert np.abs(a2.cpu().numpy() - a).mean() > 0.1 * std.min()            # fresh noise per call
