"""Oracle: policy / value networks of the PPO hot path as plain functional torch-CPU code
over a {state_dict name: tensor} dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  MLPLayer / MLPBase      openrl/modules/networks/utils/mlp.py:8-46,100-176
  RNNLayer (GRU)          openrl/modules/networks/utils/rnn.py:5-99
  Categorical/DiagGaussian openrl/modules/networks/utils/distributions.py:16-98
  ACTLayer                openrl/modules/networks/utils/act.py:45-83,160-168
  PolicyNetwork           openrl/modules/networks/policy_network.py:33-203
  ValueNetwork            openrl/modules/networks/value_network.py:33-136
Parameter names equal the reference's state_dict keys so golden
checkpoints load directly.  Parameter creation replays the reference's consumption of the
global torch generator (default nn.Linear/nn.GRU init first, then orthogonal_/constant_).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

ACT_NAMES = ["tanh", "relu", "leaky_relu", "selu"]


def activation(x, activation_id):
    return [torch.tanh, torch.relu, F.leaky_relu, F.elu][activation_id](x)


def _linear(params, name, n_in, n_out, gain, use_orthogonal=True):
    lin = nn.Linear(n_in, n_out)  # consumes the generator exactly like the reference's construction
    (nn.init.orthogonal_ if use_orthogonal else nn.init.xavier_uniform_)(lin.weight.data, gain=gain)
    nn.init.constant_(lin.bias.data, 0)
    params[name + ".weight"] = lin.weight.data.clone()
    params[name + ".bias"] = lin.bias.data.clone()


def _layernorm(params, name, n):
    params[name + ".weight"] = torch.ones(n)
    params[name + ".bias"] = torch.zeros(n)


def init_mlp_base(params, prefix, obs_dim, hidden, layer_N, activation_id, use_feature_normalization=False):
    """MLPBase.__init__ (mlp.py:100-158) -> MLPLayer.__init__ (mlp.py:8-39)."""
    gain = nn.init.calculate_gain(ACT_NAMES[activation_id])
    if use_feature_normalization:
        _layernorm(params, prefix + ".feature_norm", obs_dim)
    _linear(params, prefix + ".mlp.fc1.0", obs_dim, hidden, gain)
    _layernorm(params, prefix + ".mlp.fc1.2", hidden)
    if layer_N > 1:
        # fc_h is created once and deep-copied layer_N-1 times (mlp.py:28-34): clones share values
        tmp = {}
        _linear(tmp, "h.0", hidden, hidden, gain)
        _layernorm(tmp, "h.2", hidden)
        for k, v in tmp.items():
            params[prefix + ".mlp.fc_h." + k[2:]] = v.clone()
        for i in range(layer_N - 1):
            for k, v in tmp.items():
                params[prefix + f".mlp.fc2.{i}." + k[2:]] = v.clone()
    _linear(params, prefix + ".mlp.fc3.0", hidden, hidden, gain)
    _layernorm(params, prefix + ".mlp.fc3.1", hidden)


def init_rnn(params, prefix, hidden, recurrent_N=1):
    """RNNLayer.__init__ (rnn.py:5-37): nn.GRU default init, then biases 0 / weights orthogonal."""
    gru = nn.GRU(hidden, hidden, num_layers=recurrent_N)
    for name, p in gru.named_parameters():
        if "bias" in name:
            nn.init.constant_(p, 0)
        elif "weight" in name:
            nn.init.orthogonal_(p)
    for name, p in gru.named_parameters():
        params[prefix + ".rnn." + name] = p.data.clone()
    _layernorm(params, prefix + ".norm", hidden)


def init_policy(cfg, obs_dim, act_kind, act_dim):
    """PolicyNetwork.__init__ (policy_network.py:33-127).  act_kind: 'Discrete' | 'Box'."""
    p = {}
    init_mlp_base(p, "base", obs_dim, cfg.hidden_size, cfg.layer_N, cfg.activation_id,
                  cfg.use_feature_normalization)
    if cfg.use_recurrent_policy or getattr(cfg, "use_naive_recurrent_policy", False):
        init_rnn(p, "rnn", cfg.hidden_size, cfg.recurrent_N)
    if act_kind == "Discrete":
        _linear(p, "act.action_out.linear", cfg.hidden_size, act_dim, cfg.gain)
    else:
        _linear(p, "act.action_out.fc_mean", cfg.hidden_size, act_dim, cfg.gain)
        p["act.action_out.logstd._bias"] = torch.zeros(act_dim, 1)
    return p


def init_critic(cfg, obs_dim):
    """ValueNetwork.__init__ (value_network.py:33-111); v_out gain 1."""
    p = {}
    init_mlp_base(p, "base", obs_dim, cfg.hidden_size, cfg.layer_N, cfg.activation_id,
                  cfg.use_feature_normalization)
    if cfg.use_recurrent_policy or getattr(cfg, "use_naive_recurrent_policy", False):
        init_rnn(p, "rnn", cfg.hidden_size, cfg.recurrent_N)
    if getattr(cfg, "use_popart", False):
        # v_out = init_(PopArt(H, 1)) (value_network.py:106-109): PopArt.reset_parameters draws kaiming-uniform weights
        # and a uniform bias first (popart.py:47-55), then init_ overwrites them with orthogonal / zero
        import math

        w = torch.empty(1, cfg.hidden_size)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        b = torch.empty(1)
        nn.init.uniform_(b, -1 / math.sqrt(cfg.hidden_size), 1 / math.sqrt(cfg.hidden_size))
        (nn.init.orthogonal_ if cfg.use_orthogonal else nn.init.xavier_uniform_)(w, gain=1.0)
        p["v_out.weight"], p["v_out.bias"] = w, torch.zeros(1)
    else:
        _linear(p, "v_out", cfg.hidden_size, 1, 1.0)
    return p


def init_policy_value(cfg, obs_dim, act_kind, act_dim):
    """PolicyValueNetwork.__init__ (policy_value_network.py:33-104), cfg.use_share_model: obs_prep MLPBase ->
    common MLPLayer(H, H, layer_N=0) -> [rnn] -> v_out (gain 1) -> act; `critic_obs_prep` is the same module object as
    `obs_prep` (its state_dict entries are aliases, not new parameters)."""
    p = {}
    H = cfg.hidden_size
    init_mlp_base(p, "obs_prep", obs_dim, H, cfg.layer_N, cfg.activation_id, cfg.use_feature_normalization)
    gain = nn.init.calculate_gain(ACT_NAMES[cfg.activation_id])
    _linear(p, "common.fc1.0", H, H, gain, cfg.use_orthogonal)
    _layernorm(p, "common.fc1.2", H)
    _linear(p, "common.fc3.0", H, H, gain, cfg.use_orthogonal)
    _layernorm(p, "common.fc3.1", H)
    if cfg.use_recurrent_policy:
        init_rnn(p, "rnn", H, cfg.recurrent_N)
    _linear(p, "v_out", H, 1, 1.0, cfg.use_orthogonal)
    if act_kind == "Discrete":
        _linear(p, "act.action_out.linear", H, act_dim, cfg.gain, cfg.use_orthogonal)
    else:
        _linear(p, "act.action_out.fc_mean", H, act_dim, cfg.gain, cfg.use_orthogonal)
        p["act.action_out.logstd._bias"] = torch.zeros(act_dim, 1)
    return p


def is_shared(p):
    return "common.fc1.0.weight" in p


def shared_trunk(p, cfg, obs):
    """obs_prep -> common (policy_value_network.py:117-124,152-156,166-172): the feature both heads read."""
    x = mlp_base(p, "obs_prep", obs, cfg.layer_N, cfg.activation_id)
    h = activation(F.linear(x, p["common.fc1.0.weight"], p["common.fc1.0.bias"]), cfg.activation_id)
    h = F.layer_norm(h, h.shape[-1:], p["common.fc1.2.weight"], p["common.fc1.2.bias"])
    h = F.linear(h, p["common.fc3.0.weight"], p["common.fc3.0.bias"])
    return F.layer_norm(h, h.shape[-1:], p["common.fc3.1.weight"], p["common.fc3.1.bias"])


def mlp_base(p, prefix, x, layer_N, activation_id):
    if prefix + ".feature_norm.weight" in p:
        x = F.layer_norm(x, x.shape[-1:], p[prefix + ".feature_norm.weight"], p[prefix + ".feature_norm.bias"])
    h = F.linear(x, p[prefix + ".mlp.fc1.0.weight"], p[prefix + ".mlp.fc1.0.bias"])
    h = activation(h, activation_id)
    h = F.layer_norm(h, h.shape[-1:], p[prefix + ".mlp.fc1.2.weight"], p[prefix + ".mlp.fc1.2.bias"])
    for i in range(layer_N - 1):
        h = F.linear(h, p[prefix + f".mlp.fc2.{i}.0.weight"], p[prefix + f".mlp.fc2.{i}.0.bias"])
        h = activation(h, activation_id)
        h = F.layer_norm(h, h.shape[-1:], p[prefix + f".mlp.fc2.{i}.2.weight"], p[prefix + f".mlp.fc2.{i}.2.bias"])
    h = F.linear(h, p[prefix + ".mlp.fc3.0.weight"], p[prefix + ".mlp.fc3.0.bias"])
    h = F.layer_norm(h, h.shape[-1:], p[prefix + ".mlp.fc3.1.weight"], p[prefix + ".mlp.fc3.1.bias"])
    return h


def gru_cell(p, prefix, x, h):
    """One GRU step (torch.nn.GRU equations, gate order r,z,n)."""
    w_ih, w_hh = p[prefix + ".rnn.weight_ih_l0"], p[prefix + ".rnn.weight_hh_l0"]
    b_ih, b_hh = p[prefix + ".rnn.bias_ih_l0"], p[prefix + ".rnn.bias_hh_l0"]
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h


def rnn_layer(p, prefix, x, hxs, masks):
    """RNNLayer.forward (rnn.py:39-99) restated as "h <- GRU(x_t, h * mask_t)" per step
    (SURVEY.md §7 trap (ii)).  x: (L*n, H) time-major or (n, H); hxs: (n, 1, H); masks (L*n, 1)."""
    n = hxs.shape[0]
    H = x.shape[-1]
    L = x.shape[0] // n
    xs = x.view(L, n, H)
    ms = masks.view(L, n, 1)
    h = hxs[:, 0]
    outs = []
    for t in range(L):
        h = gru_cell(p, prefix, xs[t], h * ms[t])
        outs.append(h)
    out = torch.stack(outs).reshape(L * n, H)
    out = F.layer_norm(out, (H,), p[prefix + ".norm.weight"], p[prefix + ".norm.bias"])
    return out, h.unsqueeze(1)


def is_recurrent(cfg):
    """`_use_naive_recurrent_policy or _use_recurrent_policy` (policy_network.py:88-97, value_network.py:88-97)."""
    return bool(cfg.use_recurrent_policy or getattr(cfg, "use_naive_recurrent_policy", False))


def policy_features(p, cfg, obs, rnn_states=None, masks=None):
    f = shared_trunk(p, cfg, obs) if is_shared(p) else mlp_base(p, "base", obs, cfg.layer_N, cfg.activation_id)
    if is_recurrent(cfg):
        f, rnn_states = rnn_layer(p, "rnn", f, rnn_states, masks)
    return f, rnn_states


def categorical_logits(p, feat, action_masks=None):
    x = F.linear(feat, p["act.action_out.linear.weight"], p["act.action_out.linear.bias"])
    if action_masks is not None:
        x = x.masked_fill(action_masks == 0, -6e4)  # distributions.py:71 (in-place there: same values/grads)
    return x - x.logsumexp(dim=-1, keepdim=True)  # torch.distributions.Categorical(logits=...)


def sample_categorical(norm_logits, exp_noise):
    """torch.multinomial(probs, 1) for one sample == argmax(probs / q), q ~ Exp(1) drawn by
    `empty_like(probs).exponential_(1)` from the same generator."""
    probs = F.softmax(norm_logits, dim=-1)
    return torch.argmax(probs / exp_noise, dim=-1, keepdim=True)


def policy_act(p, cfg, obs, action_masks=None, rnn_states=None, masks=None, deterministic=False, exp_noise=None):
    """PolicyNetwork.forward_original (policy_network.py:130-162), Discrete head.
    exp_noise=None draws it from the global generator (as the reference does)."""
    feat, rnn_states = policy_features(p, cfg, obs, rnn_states, masks)
    nl = categorical_logits(p, feat, action_masks)
    if deterministic:
        actions = F.softmax(nl, -1).argmax(dim=-1, keepdim=True)
    else:
        if exp_noise is None:
            exp_noise = torch.empty_like(nl).exponential_(1)
        actions = sample_categorical(nl, exp_noise)
    logp = nl.gather(-1, actions)
    return actions, logp, rnn_states


def policy_eval(p, cfg, obs, actions, action_masks=None, active_masks=None, rnn_states=None, masks=None):
    """PolicyNetwork.eval_actions (policy_network.py:164-203) + ACTLayer.evaluate_actions
    (act.py:160-168): log-probs (B,1) and masked-mean entropy."""
    feat, _ = policy_features(p, cfg, obs, rnn_states, masks)
    nl = categorical_logits(p, feat, action_masks)
    logp = nl.gather(-1, actions.long())
    min_real = torch.finfo(nl.dtype).min
    ent = -(torch.clamp(nl, min=min_real) * F.softmax(nl, -1)).sum(-1)
    if active_masks is not None and cfg.use_policy_active_masks:
        ent = (ent * active_masks.squeeze(-1)).sum() / active_masks.sum()
    else:
        ent = ent.mean()
    return logp, ent


def gaussian_params(p, feat):
    """DiagGaussian.forward (distributions.py:75-98): mean = fc_mean(x), std = exp(logstd bias)."""
    mean = F.linear(feat, p["act.action_out.fc_mean.weight"], p["act.action_out.fc_mean.bias"])
    logstd = torch.zeros_like(mean) + p["act.action_out.logstd._bias"].t().view(1, -1)
    return mean, logstd.exp()


def policy_act_gaussian(p, cfg, obs, deterministic=False, normal_noise=None):
    """ACTLayer.forward, continuous branch (act.py:74-77): Normal.sample() == torch.normal(mean, std)
    == N(0,1) noise * std + mean with the noise drawn by `empty.normal_()` from the global generator;
    log-probs stay per dimension (distributions.py:35-37)."""
    feat, _ = policy_features(p, cfg, obs)
    mean, std = gaussian_params(p, feat)
    if deterministic:
        actions = mean
    else:
        if normal_noise is None:
            actions = torch.normal(mean, std)
        else:
            actions = normal_noise * std + mean
    logp = torch.distributions.Normal(mean, std).log_prob(actions)
    return actions, logp


def policy_eval_gaussian(p, cfg, obs, actions, active_masks=None):
    """ACTLayer.evaluate_actions, continuous branch (act.py:150-158)."""
    feat, _ = policy_features(p, cfg, obs)
    mean, std = gaussian_params(p, feat)
    dist = torch.distributions.Normal(mean, std)
    logp = dist.log_prob(actions)
    ent = dist.entropy()
    if active_masks is not None and cfg.use_policy_active_masks:
        ent = (ent * active_masks).sum() / active_masks.sum()
    else:
        ent = ent.mean()
    return logp, ent


def critic_forward(p, cfg, obs, rnn_states=None, masks=None):
    """ValueNetwork.forward (value_network.py:113-136); PolicyValueNetwork.get_values for a shared model."""
    f = shared_trunk(p, cfg, obs) if is_shared(p) else mlp_base(p, "base", obs, cfg.layer_N, cfg.activation_id)
    if is_recurrent(cfg):
        f, rnn_states = rnn_layer(p, "rnn", f, rnn_states, masks)
    return F.linear(f, p["v_out.weight"], p["v_out.bias"]), rnn_states
