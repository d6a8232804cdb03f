"""Oracle: GAE / returns and advantage normalisation (numpy float32).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  ReplayData.compute_returns      openrl/buffers/replay_data.py:320-423
  ValueNorm.running_mean_var      openrl/modules/utils/valuenorm.py:51-57
  ValueNorm.denormalize           openrl/modules/utils/valuenorm.py:92-106
  PPOAlgorithm.train_ppo (adv)    openrl/algorithms/ppo.py:384-409
Arrays keep the reference's shapes: rewards (T,N,A,1), value_preds/masks/bad_masks/returns
(T+1,N,A,1).  gamma / gae_lambda are Python floats, exactly as in the reference, so numpy's
weak-scalar promotion reproduces its float32 rounding sequence.
"""
import numpy as np


def vn_mean_var(vn_state):
    """(running_mean, running_mean_sq, debiasing_term) -> (mean, var), float32."""
    rm, rms, db = (np.float32(x) for x in vn_state)
    d = np.maximum(db, np.float32(1e-5))
    mean = rm / d
    mean_sq = rms / d
    var = np.maximum(mean_sq - mean * mean, np.float32(1e-2))
    return np.float32(mean), np.float32(var)


def denormalize(x, vn_state):
    mean, var = vn_mean_var(vn_state)
    return (x * np.sqrt(var) + mean).astype(np.float32)


def compute_returns(rewards, value_preds, masks, bad_masks, next_value, gamma, gae_lambda,
                    use_gae=True, use_proper_time_limits=False, vn_state=None):
    """Returns (returns, value_preds) as new arrays; vn_state=None means "no normaliser"."""
    rewards = rewards.astype(np.float32)
    value_preds = value_preds.astype(np.float32).copy()
    masks = masks.astype(np.float32)
    bad_masks = None if bad_masks is None else bad_masks.astype(np.float32)
    T = rewards.shape[0]
    returns = np.zeros_like(value_preds)
    dn = (lambda x: denormalize(x, vn_state)) if vn_state is not None else (lambda x: x)
    if use_proper_time_limits:
        if use_gae:
            value_preds[-1] = next_value
            gae = 0
            for step in reversed(range(T)):
                delta = rewards[step] + gamma * dn(value_preds[step + 1]) * masks[step + 1] - dn(value_preds[step])
                if vn_state is not None:
                    gae = delta + gamma * gae_lambda * gae * masks[step + 1]
                else:
                    gae = delta + gamma * gae_lambda * masks[step + 1] * gae
                gae = gae * bad_masks[step + 1]
                returns[step] = gae + dn(value_preds[step])
        else:
            returns[-1] = next_value
            for step in reversed(range(T)):
                returns[step] = (
                    returns[step + 1] * gamma * masks[step + 1] + rewards[step]
                ) * bad_masks[step + 1] + (1 - bad_masks[step + 1]) * dn(value_preds[step])
    else:
        if use_gae:
            value_preds[-1] = next_value
            gae = 0
            for step in reversed(range(T)):
                delta = rewards[step] + gamma * dn(value_preds[step + 1]) * masks[step + 1] - dn(value_preds[step])
                gae = delta + gamma * gae_lambda * masks[step + 1] * gae
                returns[step] = gae + dn(value_preds[step])
        else:
            returns[-1] = next_value
            for step in reversed(range(T)):
                returns[step] = returns[step + 1] * gamma * masks[step + 1] + rewards[step]
    return returns, value_preds


def advantages(returns, value_preds, active_masks, vn_state=None, use_adv_normalize=False):
    """ppo.py:384-409 — raw advantages and the (always applied) masked normalisation."""
    if vn_state is not None:
        adv = returns[:-1] - denormalize(value_preds[:-1], vn_state)
    else:
        adv = returns[:-1] - value_preds[:-1]
    raw = adv.copy()
    if use_adv_normalize:
        adv = (adv - adv.mean()) / (adv.std() + 1e-5)
    adv_copy = adv.copy()
    adv_copy[active_masks[:-1] == 0.0] = np.nan
    mean = np.nanmean(adv_copy)
    std = np.nanstd(adv_copy)
    adv = (adv - mean) / (std + 1e-5)
    return raw.astype(np.float32), adv.astype(np.float32)
