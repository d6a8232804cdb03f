"""CPU emulation of the DEVICE recurrent update pipeline against the unmodified reference's trace
(tests/golden/trace_cartpole_gru.npz: single agent, data_chunk_length 4, two minibatches per epoch, ValueNorm,
episodes ending inside chunks).

Everything the device does is replayed with the same building blocks and the same assembly, on the CPU:
  chunk -> buffer-row mapping            openrl_b200.buffers.replay_data.chunk_row_indices (the product helper)
  chunked forward / BPTT, per-row tape   openrl_b200/csrc/orl_rnn_core.h compiled by g++ (the sequential statement of the
                                         step that orl_rnn_warp.cuh runs warp-cooperatively and is checked against)
  per-step losses and dL/dout            numpy restatement of orl_loss.cuh (pg_term, value_term) as used in
                                         rnn_chunk_warp_kernel: active-mask weights 1/sum(active), entropy bonus
  dW = sum P^T Q tape reductions         the job table of orl_rnn.cu::make_jobs
  clip + Adam + ValueNorm commit         rnn_apply_kernel
and the six scalars of every update plus the parameters after the iteration must equal the reference's."""
import ctypes
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import loop, nets, ppo
from test_rnn_core_cpu import _ptr, shim  # noqa: F401  (pytest fixture)

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


def test_device_recurrent_pipeline_on_cpu_matches_reference_trace(shim):  # noqa: F811
    from openrl_b200.buffers.replay_data import chunk_row_indices

    d = np.load(os.path.join(GOLDEN, "trace_cartpole_gru.npz"), allow_pickle=True)
    cfg = loop.cfg_from_flags(str(d["meta/flags"]))
    N, T, L, dim, n = int(d["meta/env_num"]), cfg.episode_length, cfg.data_chunk_length, 4, 2
    B = N
    torch.manual_seed(0)
    order_p = list(nets.init_policy(cfg, dim, "Discrete", n).keys())
    order_c = list(nets.init_critic(cfg, dim).keys())
    Pp = np.concatenate([d[f"init/policy.{k}"].reshape(-1) for k in order_p]).astype(np.float64)
    Pc = np.concatenate([d[f"init/critic.{k}"].reshape(-1) for k in order_c]).astype(np.float64)
    assert shim.shim_param_count(dim, n) == Pp.size and shim.shim_param_count(dim, 1) == Pc.size
    opt_p, opt_c = _Adam(Pp.size, cfg.lr, cfg.opti_eps), _Adam(Pc.size, cfg.critic_lr, cfg.opti_eps)
    vn = ppo.ValueNormState()
    width = shim.shim_tape_width()

    def run_net(P, n_out, X, H0, M, dl):
        rows = X.shape[0]
        out = np.zeros((rows, n_out), np.float32)
        tape = np.zeros((rows, width), np.float32)
        Pf = P.astype(np.float32)
        shim.shim_chunk_fwdbwd(_ptr(Pf), dim, n_out, cfg.activation_id, L, rows // L, _ptr(X), _ptr(H0), _ptr(M),
                               _ptr(np.ascontiguousarray(dl, dtype=np.float32)), _ptr(out), _ptr(tape))
        return out.astype(np.float64), tape

    for it in range(int(d["meta/iters"])):
        g = lambda k: d[f"it{it}/{k}"]  # noqa: E731
        flat = lambda a: a.reshape(a.shape[0], B, -1)  # noqa: E731
        obs, masks, active = flat(g("policy_obs")), flat(g("masks")), flat(g("active_masks"))
        hs, hc = flat(g("rnn_states")), flat(g("rnn_states_critic"))
        actions, old_lp, vpred, ret, adv = flat(g("actions")), flat(g("action_log_probs")), flat(g("value_preds")), flat(g("returns")), flat(g("advantages"))
        np.testing.assert_allclose(vn.state(), g("vn_before_update"), rtol=1e-6, atol=1e-7)
        chunks = (T * B) // L
        mbc = chunks // cfg.num_mini_batch
        got = []
        for e in range(cfg.ppo_epoch):
            perm = torch.from_numpy(g("perms")[e])
            for i in range(cfg.num_mini_batch):
                ids = perm[i * mbc:(i + 1) * mbc]
                bi = chunk_row_indices(ids, L, T, B).numpy().reshape(mbc, L)       # (chunk, step) -> buffer row t*B + row
                tm = bi.T.reshape(-1)                                              # the shim wants time-major rows l*C + c
                t_idx, r_idx = tm // B, tm % B
                X = np.ascontiguousarray(obs[t_idx, r_idx], dtype=np.float32)
                M = np.ascontiguousarray(masks[t_idx, r_idx, 0], dtype=np.float32)
                H0p = np.ascontiguousarray(hs[bi[:, 0] // B, bi[:, 0] % B], dtype=np.float32)
                H0c = np.ascontiguousarray(hc[bi[:, 0] // B, bi[:, 0] % B], dtype=np.float32)
                act_m = active[t_idx, r_idx, 0].astype(np.float64)
                w = act_m / act_m.sum()                                            # use_*_active_masks: active / sum(active)
                rows = tm.size
                # ---- policy: forward, loss, dL/dlogits, backward ----
                logits, _ = run_net(Pp, n, X, H0p, M, np.zeros((rows, n)))
                nl = logits - np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1, keepdims=True)) - logits.max(1, keepdims=True)
                pr = np.exp(nl)
                a_idx = actions[t_idx, r_idx, 0].astype(int)
                lp = nl[np.arange(rows), a_idx]
                ratio = np.exp(lp - old_lp[t_idx, r_idx, 0])
                A = adv[t_idx, r_idx, 0].astype(np.float64)
                s1, s2 = ratio * A, np.clip(ratio, 1 - cfg.clip_param, 1 + cfg.clip_param) * A
                inside = (ratio >= 1 - cfg.clip_param) & (ratio <= 1 + cfg.clip_param)
                sel = np.where(s1 < s2, 1.0, np.where(s1 > s2, 0.0, np.where(inside, 1.0, 0.5)))
                ent = -(pr * nl).sum(1)
                policy_loss, entropy = float((-np.minimum(s1, s2) * w).sum()), float((ent * w).sum())
                dlp = (-sel * A * ratio) * w
                onehot = np.eye(n)[a_idx]
                dl = dlp[:, None] * (onehot - pr) + (cfg.entropy_coef * w)[:, None] * pr * (nl + ent[:, None])
                _, tape = run_net(Pp, n, X, H0p, M, dl)
                gp = _tape_to_grads(tape, dim, n)
                # ---- critic: ValueNorm update with this minibatch's returns, clipped Huber loss ----
                rb = ret[t_idx, r_idx, 0].astype(np.float64)
                vn.update(torch.from_numpy(rb.astype(np.float32)).view(-1, 1))
                target = vn.normalize(torch.from_numpy(rb.astype(np.float32)).view(-1, 1)).numpy()[:, 0].astype(np.float64)
                values, _ = run_net(Pc, 1, X, H0c, M, np.zeros((rows, 1)))
                vl, dv = _value_term(values[:, 0], vpred[t_idx, r_idx, 0].astype(np.float64), target, cfg.clip_param, cfg.huber_delta)
                value_loss = float((vl * w).sum())
                _, tape = run_net(Pc, 1, X, H0c, M, (cfg.value_loss_coef * w * dv)[:, None])
                gc = _tape_to_grads(tape, dim, 1)
                # ---- clip + Adam ----
                agn, cgn = float(np.sqrt((gp * gp).sum())), float(np.sqrt((gc * gc).sum()))
                Pp = opt_p.step(Pp, gp * min(cfg.max_grad_norm / (agn + 1e-6), 1.0))
                Pc = opt_c.step(Pc, gc * min(cfg.max_grad_norm / (cgn + 1e-6), 1.0))
                got.append([value_loss, cgn, policy_loss, entropy, agn, float(ratio.mean())])
        np.testing.assert_allclose(np.array(got), g("updates"), rtol=3e-4, atol=3e-6)
        off = 0
        for k in order_p:
            want = g(f"params/policy.{k}")
            np.testing.assert_allclose(Pp[off:off + want.size].reshape(want.shape), want, rtol=2e-3, atol=1e-5, err_msg=k)
            off += want.size
        off = 0
        for k in order_c:
            want = g(f"params/critic.{k}")
            np.testing.assert_allclose(Pc[off:off + want.size].reshape(want.shape), want, rtol=2e-3, atol=1e-5, err_msg=k)
            off += want.size
        np.testing.assert_allclose(vn.state(), g("vn_after_update"), rtol=1e-5, atol=1e-7)
