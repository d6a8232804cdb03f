"""Hardware check of the env-sharded multi-GPU path (run under torchrun, one rank per GPU):

  (1) ONE iteration (rollout + critic + GAE + 4 updates) of G global envs sharded over W ranks must leave every rank
      with the parameters a single process gets on the unsharded G-env run: trajectories bit-identical (same per-env
      PCG64 streams, same Philox noise per global row), parameters within 1e-5 (+ the Adam noise-floor allowance);
  (2) all ranks hold identical parameters after k further iterations (replicas stay in lockstep).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
Writes gpurun_out/multi_gpu_check.json on rank 0 and exits non-zero on a mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import bench
from openrl_b200 import parallel


def run(rank, world, envs, offset_rank, iters, flags):
    bench.WORKLOADS["chk"] = dict(env="CartPole-v1", envs=envs, T=32, flags=flags)
    cfg, env, net, agent = bench.build_agent(offset_rank, world, "chk")
    drv = bench.make_driver(cfg, env, net, agent, rank, world)
    outs = []
    for _ in range(iters):
        drv._rollout_launch()
        drv.compute_returns()
        b = drv.buffer.data
        snap = {k: getattr(b, k).clone() for k in ("actions", "policy_obs", "rewards", "masks", "returns")}
        drv.trainer.train_async(b)
        info = drv.trainer.read_train_info()
        b.after_update()
        params = torch.cat([net.module.models[m].flat_params.clone() for m in ("policy", "critic")])
        outs.append((snap, info, params))
    run.exchange = ("NVLink peer-memory sum inside orl_ppo_apply_peer" if getattr(drv.trainer, "peer", None) is not None
                    else "all-reduce (NCCL / symmetric-memory one-shot)") if world > 1 else "none"
    return cfg, outs


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    G = 512
    per = G // world
    flags = ["--seed", "0", "--episode_length", "32", "--ppo_epoch", "4", "--num_mini_batch", "1", "--log_interval", "1000000"]
    iters = 4
    cfg, sharded = run(rank, world, per, rank, iters, flags)
    exchange = run.exchange
    # (2) lockstep: every rank's parameters equal rank 0's, exactly (same all-reduced bucket, same Adam)
    p_last = sharded[-1][2]
    ref = p_last.clone()
    dist.broadcast(ref, 0)
    lock = torch.tensor([float((p_last - ref).abs().max())], device="cuda")
    dist.all_reduce(lock, op=dist.ReduceOp.MAX)
    # gather iteration-0 snapshots of every rank on rank 0
    keys = ("actions", "policy_obs", "rewards", "masks", "returns")
    gathered = {}
    for k in keys:
        t = sharded[0][0][k].contiguous()
        lst = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, lst, dst=0)
        if rank == 0:
            gathered[k] = torch.cat(lst, dim=1)   # env axis
    dist.barrier()
    ok, report = True, {}
    if rank == 0:
        parallel.force_single_process(True)
        _, single = run(0, 1, G, 0, iters, flags)
        parallel.force_single_process(False)
        for k in keys:
            same = bool(torch.equal(gathered[k], single[0][0][k]))
            report[f"it0_{k}_bit_identical"] = same
            ok &= same
        d0 = (sharded[0][2] - single[0][2]).abs()
        lr = cfg.lr
        report["it0_params_max_abs_diff"] = float(d0.max())
        report["it0_params_frac_within_1e-5"] = float((d0 <= 1e-5).float().mean())
        report["it0_info_sharded"] = sharded[0][1]
        report["it0_info_single"] = single[0][1]
        rel = max(abs(sharded[0][1][k] - single[0][1][k]) / max(abs(single[0][1][k]), 1e-12) for k in single[0][1])
        report["it0_info_max_rel_diff"] = rel
        ok &= rel < 1e-4 and float(d0.max()) <= 0.1 * lr and float((d0 <= 1e-5).float().mean()) > 0.995
        report["later_iterations_params_max_abs_diff"] = [float((sharded[i][2] - single[i][2]).abs().max()) for i in range(1, iters)]
        report["lockstep_max_abs_diff_over_ranks"] = float(lock.item())
        ok &= float(lock.item()) == 0.0
        report.update(exchange=exchange, world=world, global_envs=G, T=32, epochs=4, ok=bool(ok))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_check_{world}.json"), "w") as f:
            json.dump(report, f, indent=1)
        print(json.dumps(report))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
