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


def _global_targets(stats):
    """Value targets from the GLOBAL minibatch moments, as the device forms them (orl_ppo.cu vn_updated + vn_mean_std,
    valuenorm.py:59-90) from a zero ValueNorm state: stats = [sum ret, sum ret^2, sum active, rows]."""
    beta = 0.99999
    bm, bsq = stats[0] / stats[3], stats[1] / stats[3]
    rm, rmsq, deb = (1 - beta) * bm, (1 - beta) * bsq, (1 - beta)
    mean = rm / max(deb, 1e-5)
    var = max(rmsq / max(deb, 1e-5) - mean * mean, 1e-2)
    return lambda rb: (rb - mean) / np.sqrt(var)


def _recurrent_setup(so_path):
    import ctypes

    from conftest import GOLDEN
    from oracle import loop, nets
    import rnn_pipeline_helpers as hp

    d = np.load(os.path.join(GOLDEN, "trace_cartpole_gru.npz"), allow_pickle=True)
    cfg = loop.cfg_from_flags(str(d["meta/flags"]))
    dim, n, B = 4, 2, int(d["meta/env_num"])
    torch.manual_seed(0)
    order_p, order_c = list(nets.init_policy(cfg, dim, "Discrete", n).keys()), list(nets.init_critic(cfg, dim).keys())
    Pp = np.concatenate([d[f"init/policy.{k}"].reshape(-1) for k in order_p]).astype(np.float64)
    Pc = np.concatenate([d[f"init/critic.{k}"].reshape(-1) for k in order_c]).astype(np.float64)
    buf = hp.load_trace_buffers(d, 0, B)
    ids = d["it0/perms"][0][:32]                       # the first minibatch of the trace: 32 chunks of 4 steps
    return hp, ctypes.CDLL(so_path), cfg, Pp, Pc, buf, ids, dim, n


def _local_stats(cfg, buf, ids):
    from openrl_b200.buffers.replay_data import chunk_row_indices

    B = buf["masks"].shape[1]
    bi = chunk_row_indices(torch.as_tensor(ids), cfg.data_chunk_length, cfg.episode_length, B).numpy()
    r = buf["ret"][bi // B, bi % B, 0].astype(np.float64)
    return np.array([r.sum(), (r * r).sum(), buf["active"][bi // B, bi % B, 0].sum(), float(bi.size)])


def _recurrent_worker(rank, world, port, so_path, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openrl_b200 import parallel

    hp, shim, cfg, Pp, Pc, buf, ids, dim, n = _recurrent_setup(so_path)
    mine = ids[rank::world] if rank == 0 else ids[rank::world][:-3]   # uneven shards (16 and 13 chunks)
    stats = torch.from_numpy(_local_stats(cfg, buf, mine))
    parallel.allreduce_sum_(stats)                                     # the minibatch-moments all-reduce
    gp, gc, sums, _ = hp.minibatch_buckets(shim, cfg, Pp, Pc, buf, mine, dim, n, _global_targets(stats.numpy()),
                                           act_sum=float(stats[2]))
    bucket = torch.from_numpy(np.concatenate([gp, gc, sums]))
    parallel.allreduce_sum_(bucket)                                    # the single gradient-bucket all-reduce
    q.put((rank, stats.numpy(), bucket.numpy(), mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_recurrent_bucket_equals_global_minibatch(tmp_path):
    """The recurrent update's multi-GPU contract under gloo: every rank weights its chunks by the GLOBAL
    sum(active) and uses targets from the GLOBAL batch moments; the all-reduced bucket equals the bucket of one
    process working on the union of the shards (CPU replay of the device pipeline, see test_rnn_pipeline_cpu.py)."""
    import subprocess

    from conftest import ROOT

    so = str(tmp_path / "librnnshim.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "openrl_b200", "csrc"),
                    os.path.join(ROOT, "tests", "rnn_core_shim.cpp"), "-o", so], check=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_recurrent_worker, args=(r, 2, port, so, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hp, shim, cfg, Pp, Pc, buf, ids, dim, n = _recurrent_setup(so)
    union = np.concatenate([res[0][3], res[1][3]])
    stats = _local_stats(cfg, buf, union)
    np.testing.assert_allclose(res[0][1], stats, rtol=1e-12)
    gp, gc, sums, _ = hp.minibatch_buckets(shim, cfg, Pp, Pc, buf, union, dim, n, _global_targets(stats), act_sum=stats[2])
    want = np.concatenate([gp, gc, sums])
    for r in res:
        np.testing.assert_allclose(r[2], want, rtol=1e-6, atol=1e-9)
    assert np.abs(want[:gp.size]).max() > 1e-3


def _peer_fallback_worker(rank, world, port, q):
    import warnings

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openrl_b200 import parallel

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        # no NVLink peer mapping on a CPU / gloo group: every rank must come back with None (and say why), none may hang
        pb = parallel.PeerBucket.create(4096, 64, torch.device("cpu"))
        msgs = [str(x.message) for x in w]
    os.environ["ORL_PEER_APPLY"] = "0"
    off = parallel.PeerBucket.create(4096, 64, torch.device("cpu"))
    q.put((rank, pb is None, off is None, any("peer-memory gradient exchange unavailable" in m for m in msgs)))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_bucket_falls_back_on_every_rank_when_the_mapping_is_unavailable():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_peer_fallback_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res), res
