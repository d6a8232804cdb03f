"""The GAE oracle against golden vectors produced by the unmodified reference
(ReplayData.compute_returns, openrl/buffers/replay_data.py:320-423; 8 branches — the matrix
of the reference's tests/test_buffer/test_generator.py:28-89)."""
import glob
import os

import numpy as np
import pytest

from oracle import gae as ogae

from conftest import GOLDEN

FILES = sorted(glob.glob(os.path.join(GOLDEN, "gae_*.npz")))


def test_all_eight_branches_present():
    assert len(FILES) == 8


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_bit_exact_vs_reference(path):
    d = np.load(path)
    name = os.path.basename(path)
    use_gae, ptl, vn = (name[5] == "1"), (name[8] == "1"), (name[11] == "1")
    ret, vp = ogae.compute_returns(
        d["rewards"], d["value_preds"], d["masks"], d["bad_masks"], d["next_value"],
        float(d["gamma"]), float(d["gae_lambda"]), use_gae=use_gae, use_proper_time_limits=ptl,
        vn_state=d["vn_state"] if vn else None,
    )
    assert np.array_equal(ret, d["returns"])
    assert np.array_equal(vp, d["value_preds_after"])


def test_advantages_match_reference_trace():
    d = np.load(os.path.join(GOLDEN, "trace_cartpole.npz"), allow_pickle=True)
    raw, adv = ogae.advantages(d["it0/returns"], d["it0/value_preds"], d["it0/active_masks"],
                               vn_state=d["it0/vn_before_update"])
    np.testing.assert_allclose(adv, d["it0/advantages"], rtol=0, atol=1e-6)


def test_returns_match_reference_trace():
    d = np.load(os.path.join(GOLDEN, "trace_cartpole.npz"), allow_pickle=True)
    vp = d["it1/value_preds"]
    ret, _ = ogae.compute_returns(d["it1/rewards"], vp, d["it1/masks"], d["it1/bad_masks"], vp[-1],
                                  0.99, 0.95, vn_state=d["it1/vn_before_update"])
    assert np.array_equal(ret[:-1], d["it1/returns"][:-1])
