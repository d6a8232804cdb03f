"""CUDA GAE kernel (orl_gae, through the C-ABI) against the oracle and the reference's
golden vectors.  Bar: bit-exact (float32 ops in the reference's order, no FMA contraction)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import gae as ogae

pytestmark = pytest.mark.gpu

FILES = sorted(glob.glob(os.path.join(GOLDEN, "gae_*.npz")))


def _run(orl_lib, cuda, rewards, value_preds, masks, bad_masks, active, next_value, vn_state, gamma, lam, flags,
         want_adv=True, want_stats=True):
    import torch

    from openrl_b200 import lib

    T = rewards.shape[0]
    B = int(np.prod(rewards.shape[1:]))
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)
    r, vp, m, bm, am, nv, vn = map(dev, (rewards, value_preds, masks, bad_masks, active, next_value, vn_state))
    ret = torch.zeros_like(vp)
    adv = torch.empty_like(r) if want_adv else None
    stats = torch.empty(8, dtype=torch.float64, device=cuda) if want_stats else None
    lib.check(orl_lib.orl_gae(lib.ptr(r), lib.ptr(vp), lib.ptr(m), lib.ptr(bm), lib.ptr(am), lib.ptr(nv), lib.ptr(vn),
                              lib.ptr(ret), lib.ptr(adv), lib.ptr(stats), T, B, gamma, lam, flags,
                              lib.current_stream()), "orl_gae")
    torch.cuda.synchronize()
    return (ret.cpu().numpy(), vp.cpu().numpy(), None if adv is None else adv.cpu().numpy(),
            None if stats is None else stats.cpu().numpy())


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gae_bit_exact_vs_reference_golden(orl_lib, cuda, path):
    d = np.load(path)
    name = os.path.basename(path)
    use_gae, ptl, vn = (name[5] == "1"), (name[8] == "1"), (name[11] == "1")
    flags = (1 if use_gae else 0) | (2 if ptl else 0) | (4 if vn else 0)
    ret, vp, adv, stats = _run(orl_lib, cuda, d["rewards"], d["value_preds"], d["masks"], d["bad_masks"], None,
                               d["next_value"], d["vn_state"] if vn else None, float(d["gamma"]),
                               float(d["gae_lambda"]), flags)
    want = d["returns"]
    if use_gae:
        assert np.array_equal(ret[:-1], want[:-1])
        assert np.array_equal(vp, d["value_preds_after"])
    else:
        assert np.array_equal(ret, want)


@pytest.mark.parametrize("T,N,A", [(1, 1, 1), (5, 3, 2), (128, 8, 1), (25, 64, 3), (33, 1027, 1), (16, 4096, 1)])
@pytest.mark.parametrize("flags", [1, 5, 7, 3, 0, 2, 6])
def test_gae_bit_exact_vs_oracle_random(orl_lib, cuda, T, N, A, flags):
    rng = np.random.default_rng(T * 1000 + N + flags)
    sh, sh1 = (T, N, A, 1), (T + 1, N, A, 1)
    rewards = rng.standard_normal(sh).astype(np.float32)
    vp = rng.standard_normal(sh1).astype(np.float32)
    masks = (rng.random(sh1) > 0.05).astype(np.float32)
    bad = (rng.random(sh1) > 0.1).astype(np.float32)
    active = (rng.random(sh1) > 0.2).astype(np.float32)
    nv = rng.standard_normal((N, A, 1)).astype(np.float32)
    vn_state = np.array([0.31, 2.7, 0.4], np.float32)
    use_gae, ptl, dn = bool(flags & 1), bool(flags & 2), bool(flags & 4)
    want_ret, want_vp = ogae.compute_returns(rewards, vp, masks, bad, nv, 0.99, 0.95, use_gae=use_gae,
                                             use_proper_time_limits=ptl,
                                             vn_state=vn_state if (dn and (use_gae or ptl)) else None)
    ret, vp_out, adv, stats = _run(orl_lib, cuda, rewards, vp, masks, bad, active, nv, vn_state if dn else None,
                                   0.99, 0.95, flags)
    got = ret.reshape(sh1)
    if use_gae:
        assert np.array_equal(got[:-1], want_ret[:-1])
        assert np.array_equal(vp_out.reshape(sh1), want_vp)
    else:
        assert np.array_equal(got, want_ret)
    # fused advantage: returns[:-1] - denorm(value_preds[:-1])  (ppo.py:384-399)
    vn_for_adv = vn_state if dn else None   # ppo.py:384-399 denormalises whenever a normaliser exists
    raw, _ = ogae.advantages(want_ret, want_vp if use_gae else vp, active, vn_state=vn_for_adv)
    assert np.array_equal(adv.reshape(sh), raw)
    a64 = raw.astype(np.float64)
    on = active[:-1] != 0
    np.testing.assert_allclose(stats[0], a64.sum(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(stats[1], (a64 ** 2).sum(), rtol=1e-9)
    assert stats[2] == T * N * A
    np.testing.assert_allclose(stats[3], a64[on].sum(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(stats[4], (a64[on] ** 2).sum(), rtol=1e-9)
    assert stats[7] == on.sum()
    r64 = want_ret[:-1].astype(np.float64)
    np.testing.assert_allclose(stats[5], r64.sum(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(stats[6], (r64 ** 2).sum(), rtol=1e-9)


def test_gae_large_shape_properties(orl_lib, cuda):
    """BASELINE-size property check (no oracle run): with masks == 0 everywhere the scan
    degenerates to returns[t] = rewards[t] (+0), and with gamma = 0 likewise."""
    import torch

    from openrl_b200 import lib

    T, B = 128, 1 << 18
    g = torch.Generator(device="cuda").manual_seed(0)
    r = torch.randn(T, B, device=cuda, generator=g)
    vp = torch.randn(T + 1, B, device=cuda, generator=g)
    m = torch.zeros(T + 1, B, device=cuda)
    nv = torch.randn(B, device=cuda, generator=g)
    ret = torch.empty(T + 1, B, device=cuda)
    lib.check(orl_lib.orl_gae(lib.ptr(r), lib.ptr(vp), lib.ptr(m), None, None, lib.ptr(nv), None, lib.ptr(ret), None,
                              None, T, B, 0.99, 0.95, 1, lib.current_stream()), "orl_gae")
    # delta = r - v ; gae = delta ; ret = (r - v) + v
    assert torch.equal(ret[:-1], (r - vp[:-1]) + vp[:-1])
    # vectorised (128-bit) path == scalar path on the same data, bit for bit
    m2 = (torch.rand(T + 1, B, device=cuda, generator=g) > 0.02).float()
    ret_a = torch.empty(T + 1, B, device=cuda)
    vp_a = vp.clone()
    lib.check(orl_lib.orl_gae(lib.ptr(r), lib.ptr(vp_a), lib.ptr(m2), None, None, lib.ptr(nv), None, lib.ptr(ret_a),
                              None, None, T, B, 0.99, 0.95, 1, lib.current_stream()), "orl_gae")
    # force the scalar path with a misaligned (offset by one float) copy of everything
    def off(t):
        buf = torch.empty(t.numel() + 1, device=cuda)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        return v
    r_b, vp_b, m_b, nv_b = off(r), off(vp), off(m2), off(nv)
    ret_b = off(torch.empty(T + 1, B, device=cuda))
    lib.check(orl_lib.orl_gae(lib.ptr(r_b), lib.ptr(vp_b), lib.ptr(m_b), None, None, lib.ptr(nv_b), None,
                              lib.ptr(ret_b), None, None, T, B, 0.99, 0.95, 1, lib.current_stream()), "orl_gae")
    assert torch.equal(ret_a[:-1], ret_b[:-1])
