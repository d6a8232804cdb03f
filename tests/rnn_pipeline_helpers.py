"""Numpy / g++-core building blocks of the CPU replay of the device recurrent update (see test_rnn_pipeline_cpu.py)."""
import ctypes

import numpy as np

H = 64


def _value_term(v, vp, target, clip, delta):
    """orl_loss.cuh value_term with HUBER | CLIP_VALUE (ppo.py:178-220)."""
    hub = lambda e: np.where(np.abs(e) <= delta, 0.5 * e * e, delta * (np.abs(e) - 0.5 * delta))  # noqa: E731
    hubg = lambda e: np.where(np.abs(e) <= delta, e, np.sign(e) * delta)  # noqa: E731
    diff = v - vp
    clipped = vp + np.clip(diff, -clip, clip)
    e_c, e_o = target - clipped, target - v
    l_c, l_o = hub(e_c), hub(e_o)
    inrange = (diff >= -clip) & (diff <= clip)
    dc = np.where(inrange, -hubg(e_c), 0.0)
    loss = np.maximum(l_o, l_c)
    dv = np.where(l_o > l_c, -hubg(e_o), np.where(l_c > l_o, dc, 0.5 * (-hubg(e_o)) + 0.5 * dc))
    return loss, dv


def _tape_to_grads(tape, d, n):
    """make_jobs (orl_rnn.cu): flat gradient in the reference's state_dict order."""
    t = tape.astype(np.float64)
    dz1, dz3, dgi, dgh, dlg = t[:, 0:64], t[:, 64:128], t[:, 128:320], t[:, 320:512], t[:, 512:512 + n]
    x, y1, y3, hm, o = t[:, 520:520 + d], t[:, 584:648], t[:, 648:712], t[:, 712:776], t[:, 776:840]
    parts = [dz1.T @ x, dz1.sum(0), t[:, 840:904].sum(0), t[:, 904:968].sum(0), dz3.T @ y1, dz3.sum(0), t[:, 968:1032].sum(0),
             t[:, 1032:1096].sum(0), dgi.T @ y3, dgh.T @ hm, dgi.sum(0), dgh.sum(0), t[:, 1096:1160].sum(0), t[:, 1160:1224].sum(0),
             dlg.T @ o, dlg.sum(0)]
    return np.concatenate([p.reshape(-1) for p in parts])


class _Adam:
    """torch.optim.Adam (single tensor, no amsgrad) on a flat float64 vector; rnn_apply_kernel."""

    def __init__(self, n, lr, eps):
        self.m, self.v, self.t, self.lr, self.eps = np.zeros(n), np.zeros(n), 0, lr, eps

    def step(self, p, g, b1=0.9, b2=0.999):
        self.t += 1
        self.m = self.m + (g - self.m) * (1 - b1)
        self.v = self.v * b2 + g * g * (1 - b2)
        denom = np.sqrt(self.v) / np.sqrt(1 - b2 ** self.t) + self.eps
        return p - (self.lr / (1 - b1 ** self.t)) * (self.m / denom)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_net(shim, P, dim, n_out, act_id, L, X, H0, M, dl):
    """L-step chunks through the g++ build of orl_rnn_core.h: head outputs (rows, n_out) and the per-row tape."""
    rows = X.shape[0]
    out = np.zeros((rows, n_out), np.float32)
    tape = np.zeros((rows, shim.shim_tape_width()), np.float32)
    Pf = P.astype(np.float32)
    shim.shim_chunk_fwdbwd(ptr(Pf), dim, n_out, act_id, L, rows // L, ptr(X), ptr(H0), ptr(M),
                           ptr(np.ascontiguousarray(dl, dtype=np.float32)), ptr(out), ptr(tape))
    return out.astype(np.float64), tape


def minibatch_buckets(shim, cfg, Pp, Pc, buf, ids, dim, n, target_fn, act_sum=None):
    """Gradient buckets (policy, critic) and loss sums of the chunks `ids` exactly as the device assembles them.
    `act_sum`: the GLOBAL sum of active masks (all ranks) for the 1/sum(active) weights; None = this shard's own.
    `target_fn(returns)` gives the value targets (ValueNorm-normalised returns).  `buf` holds (T+1|T, B, k) arrays."""
    import torch

    from openrl_b200.buffers.replay_data import chunk_row_indices

    T, L, B = cfg.episode_length, cfg.data_chunk_length, buf["masks"].shape[1]
    mbc = len(ids)
    bi = chunk_row_indices(torch.as_tensor(ids), L, T, B).numpy().reshape(mbc, L)   # (chunk, step) -> buffer row t*B + row
    tm = bi.T.reshape(-1)                                                           # the shim wants time-major rows l*C + c
    t_idx, r_idx = tm // B, tm % B
    X = np.ascontiguousarray(buf["obs"][t_idx, r_idx], dtype=np.float32)
    M = np.ascontiguousarray(buf["masks"][t_idx, r_idx, 0], dtype=np.float32)
    H0p = np.ascontiguousarray(buf["hs"][bi[:, 0] // B, bi[:, 0] % B], dtype=np.float32)
    H0c = np.ascontiguousarray(buf["hc"][bi[:, 0] // B, bi[:, 0] % B], dtype=np.float32)
    act_m = buf["active"][t_idx, r_idx, 0].astype(np.float64)
    w = act_m / (act_m.sum() if act_sum is None else act_sum)                       # use_*_active_masks: active / sum(active)
    rows = tm.size
    logits, _ = run_net(shim, Pp, dim, n, cfg.activation_id, L, X, H0p, M, np.zeros((rows, n)))
    mx = logits.max(1, keepdims=True)
    nl = logits - mx - np.log(np.exp(logits - mx).sum(1, keepdims=True))
    pr = np.exp(nl)
    a_idx = buf["actions"][t_idx, r_idx, 0].astype(int)
    lp = nl[np.arange(rows), a_idx]
    ratio = np.exp(lp - buf["old_lp"][t_idx, r_idx, 0])
    A = buf["adv"][t_idx, r_idx, 0].astype(np.float64)
    s1, s2 = ratio * A, np.clip(ratio, 1 - cfg.clip_param, 1 + cfg.clip_param) * A
    inside = (ratio >= 1 - cfg.clip_param) & (ratio <= 1 + cfg.clip_param)
    sel = np.where(s1 < s2, 1.0, np.where(s1 > s2, 0.0, np.where(inside, 1.0, 0.5)))
    ent = -(pr * nl).sum(1)
    dlp = (-sel * A * ratio) * w
    dl = dlp[:, None] * (np.eye(n)[a_idx] - pr) + (cfg.entropy_coef * w)[:, None] * pr * (nl + ent[:, None])
    _, tape = run_net(shim, Pp, dim, n, cfg.activation_id, L, X, H0p, M, dl)
    gp = _tape_to_grads(tape, dim, n)
    rb = buf["ret"][t_idx, r_idx, 0].astype(np.float64)
    values, _ = run_net(shim, Pc, dim, 1, cfg.activation_id, L, X, H0c, M, np.zeros((rows, 1)))
    vl, dv = _value_term(values[:, 0], buf["vpred"][t_idx, r_idx, 0].astype(np.float64), target_fn(rb), cfg.clip_param, cfg.huber_delta)
    _, tape = run_net(shim, Pc, dim, 1, cfg.activation_id, L, X, H0c, M, (cfg.value_loss_coef * w * dv)[:, None])
    gc = _tape_to_grads(tape, dim, 1)
    sums = np.array([(vl * w).sum(), (-np.minimum(s1, s2) * w).sum(), (ent * w).sum(), ratio.sum()])
    stats = np.array([rb.sum(), (rb * rb).sum(), act_m.sum(), float(rows)])          # what orl_minibatch_stats + norm_rows carry
    return gp, gc, sums, stats


def load_trace_buffers(d, it, B):
    g = lambda k: d[f"it{it}/{k}"]  # noqa: E731
    flat = lambda a: a.reshape(a.shape[0], B, -1)  # noqa: E731
    return dict(obs=flat(g("policy_obs")), masks=flat(g("masks")), active=flat(g("active_masks")), hs=flat(g("rnn_states")),
                hc=flat(g("rnn_states_critic")), actions=flat(g("actions")), old_lp=flat(g("action_log_probs")),
                vpred=flat(g("value_preds")), ret=flat(g("returns")), adv=flat(g("advantages")))
