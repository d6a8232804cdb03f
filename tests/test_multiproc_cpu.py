"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic (openrl_b200/parallel.py):
env sharding reproduces the unsharded seeding, and the two SUM all-reduces of the path give
global-batch moments / gradient buckets."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openrl_b200 import parallel
    from openrl_b200.envs.vec_env.device_venv import _pcg64_streams

    first, n = parallel.env_shard(8)
    streams = _pcg64_streams(7, n, first)
    # rollout moments: each rank holds the moments of its shard; SUM all-reduce -> global
    rng = np.random.default_rng(100 + rank)
    adv = rng.standard_normal(1000)
    stats = torch.tensor([adv.sum(), (adv ** 2).sum(), adv.size], dtype=torch.float64)
    parallel.allreduce_sum_(stats)
    bucket = torch.full((16,), float(rank + 1))
    parallel.allreduce_sum_(bucket)
    q.put((rank, first, n, streams, stats.numpy(), bucket.numpy(), adv))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    from openrl_b200.envs.vec_env.device_venv import _pcg64_streams

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = _pcg64_streams(7, 8, 0)
    got = np.concatenate([res[0][3], res[1][3]], axis=1)
    assert np.array_equal(got, full)  # sharded seeding == unsharded seeding
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 4, 4, 4)
    adv_all = np.concatenate([res[0][6], res[1][6]])
    for r in res:
        np.testing.assert_allclose(r[4], [adv_all.sum(), (adv_all ** 2).sum(), adv_all.size], rtol=1e-12)
        assert np.array_equal(r[5], np.full(16, 3.0, np.float32))


def test_single_process_is_noop():
    from openrl_b200 import parallel

    t = torch.ones(4)
    assert parallel.allreduce_sum_(t) is t and t.sum() == 4
    assert parallel.env_shard(12, 0, 1) == (0, 12)
