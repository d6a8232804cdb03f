"""PolicyNetwork (reference: openrl/modules/networks/policy_network.py:33): base -> act."""
import torch
import torch.nn as nn

from .base import ACTLayer, FlatParams, MLPBase, RNNLayer


def _policy_shape(space):
    return space["policy"].shape if space.__class__.__name__ == "Dict" else space.shape


class PolicyNetwork(nn.Module):
    def __init__(self, cfg, input_space, action_space, device=torch.device("cpu"), use_half=False, extra_args=None):
        super().__init__()
        # `_use_naive_recurrent_policy or _use_recurrent_policy` (policy_network.py:88-97): same RNNLayer either way
        self.recurrent = bool(cfg.use_recurrent_policy or cfg.use_naive_recurrent_policy)
        self.hidden_size = cfg.hidden_size
        shape = _policy_shape(input_space)
        if len(shape) != 1 or shape[0] > 64:
            raise NotImplementedError("vector observations of width <= 64 only")
        self.obs_dim = shape[0]
        self.activation_id = cfg.activation_id
        self.base = MLPBase(cfg, shape)
        if self.recurrent:   # module order base -> rnn -> act as in the reference (state_dict / flat layout)
            self.rnn = RNNLayer(self.base.output_size, self.base.output_size, cfg.recurrent_N, cfg.use_orthogonal, cfg.rnn_type)
        self.act = ACTLayer(action_space, self.base.output_size, cfg.use_orthogonal, cfg.gain)
        self.head_kind = 1 if self.act.continuous_action else 0          # lib.HEAD_GAUSSIAN / HEAD_CATEGORICAL
        self.n_actions = action_space.shape[0] if self.act.continuous_action else action_space.n  # head width
        if self.recurrent and self.act.continuous_action:
            raise NotImplementedError("recurrent policies are built for Discrete action spaces")
        if self.n_actions > 8:
            raise NotImplementedError("head widths up to 8 are built")
        self.device = torch.device(device)
        self._flat = FlatParams(self, self.device)

    @property
    def flat_params(self):
        return self._flat.flat
