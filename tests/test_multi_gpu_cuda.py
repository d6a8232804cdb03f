"""Multi-GPU hardware check (needs >= 2 GPUs on the box; skipped otherwise — the host-side sharding logic is covered on
CPU by tests/test_multiproc_cpu.py with gloo): the env-sharded run with the gradient bucket exchanged over NVLink peer
memory inside orl_ppo_reduce_peer / orl_ppo_apply_peer must reproduce the single-process run of the global batch
(trajectories bit-identical, parameters within the Adam noise floor, replicas in exact lockstep).  The reference has no
working distributed update; the oracle of this test is the unsharded device run, which the other GPU
tests pin to the reference traces."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_sharded_run_equals_unsharded_run(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, ORL_PEER_APPLY="1" if exchange == "peer" else "0")
    port = 29541 if exchange == "peer" else 29542
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                       env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    rep = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["ok"] and rep["lockstep_max_abs_diff_over_ranks"] == 0.0
    assert ("peer-memory" in rep["exchange"]) == (exchange == "peer")
